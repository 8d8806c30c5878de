#define B2S_BUILD_HASH "e6b5b39f624ebd3f"
