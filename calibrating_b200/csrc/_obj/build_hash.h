#define B2S_BUILD_HASH "c79eec3fb79428fa"
