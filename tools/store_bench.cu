// store_bench.cu — how much does the granularity of scattered stores cost on B200?
// The mask kernel writes, per 128-pixel tile, one 128-byte segment per (warp, mask) into ~160 masks that lie 305 KB
// apart.  Here: every CTA owns a window of `tile_bytes` per stream and walks the streams, writing `chunk` contiguous
// bytes per stream visit with 128-bit stores (a warp covers 512 B per instruction).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/store_bench.cu -o tools/store_bench && tools/store_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void scatter_store(float4* out, long long stream_stride_f4, int n_streams, long long tiles, int chunk_f4) {
  // tile t of the output: streams s = 0..n_streams-1 each get chunk_f4 float4 at offset t*chunk_f4
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    for (int s = threadIdx.x / 32; s < n_streams; s += blockDim.x / 32) {  // one warp per stream visit
      float4* p = out + (long long)s * stream_stride_f4 + t * chunk_f4;
      for (int i = threadIdx.x & 31; i < chunk_f4; i += 32) p[i] = v;
    }
  }
}

int main() {
  const int n_streams = 160;
  const long long stream_bytes = 76176LL * 4;            // one 276x276 fp32 mask
  const int frames = 24;                                  // independent groups of streams (like frames)
  const long long total = (long long)frames * n_streams * stream_bytes;
  float4* buf;
  cudaMalloc(&buf, total + (1 << 20));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int chunk_bytes : {128, 256, 512, 1024, 2048, 4096}) {
    const int chunk_f4 = chunk_bytes / 16;
    const long long tiles = stream_bytes / chunk_bytes;   // per frame
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      for (int f = 0; f < frames; ++f)
        scatter_store<<<148 * 2, 256>>>(buf + (long long)f * n_streams * (stream_bytes / 16), stream_bytes / 16, n_streams, tiles, chunk_f4);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("chunk %5d B per stream visit: %8.1f us  %7.0f GB/s\n", chunk_bytes, best * 1e3, frames * n_streams * tiles * (double)chunk_bytes / best / 1e6);
  }
  // contiguous fill for reference
  return 0;
}
