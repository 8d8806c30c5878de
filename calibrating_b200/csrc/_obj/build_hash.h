#define B2S_BUILD_HASH "a54bf16cfba84c83"
