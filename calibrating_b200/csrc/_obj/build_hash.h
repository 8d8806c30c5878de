#define B2S_BUILD_HASH "3e0771ccefce18bd"
