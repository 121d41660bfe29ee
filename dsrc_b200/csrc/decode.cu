// decode + synthetic generator entry points (placeholders that fail loudly until implemented)
#include "../../include/dsrc_b200.h"
extern "C" int dsrcgpu_decode_blocks(dsrcgpu_ctx*, const uint8_t*, const uint64_t*, const uint32_t*, uint32_t, uint8_t*, uint64_t, uint64_t*) { return DSRCGPU_E_UNSUPPORTED; }
extern "C" int dsrcgpu_decode_blocks_device(dsrcgpu_ctx*, const uint8_t*, const uint64_t*, const uint32_t*, uint32_t, uint8_t*, uint64_t, uint64_t*) { return DSRCGPU_E_UNSUPPORTED; }
