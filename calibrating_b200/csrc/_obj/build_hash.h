#define B2S_BUILD_HASH "13a7dfea0fc94d8e"
