// -q0 / -d0 Huffman and 2-bit paths (placeholder until the kernels land: blocks fail loudly, never silently)
#include "common.cuh"
#include "kernels.h"
__global__ void k_unsupported(Workspace ws) { u32 b = blockIdx.x * blockDim.x + threadIdx.x; if (b < ws.n_blocks && ws.state[b].status == ST_OK) ws.state[b].status = ST_UNSUPPORTED; }
void launch_q0_quality(const Workspace& ws, cudaStream_t s) { k_unsupported<<<(ws.n_blocks + 127) / 128, 128, 0, s>>>(ws); }
void launch_d0_dna(const Workspace& ws, cudaStream_t s) { k_unsupported<<<(ws.n_blocks + 127) / 128, 128, 0, s>>>(ws); }
