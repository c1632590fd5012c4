// Micro-benchmark (2 GPUs, one process): how fast can SM-issued stores fill a peer's memory over NVLink?
//   stg  : every thread stores 16 B, a warp 512 contiguous bytes, a CTA sweeps contiguous 8 KB chunks (what K3 does)
//   bulk : the CTA keeps a 64 KB tile in shared memory and pushes it with cp.async.bulk (TMA) 8 KB at a time
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_store nvlink_store.cu && ./nvlink_store
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_stg(double2 *dst, size_t elems_per_cta_iter, int iters, size_t total_elems) {
  // CTA b, iteration i writes chunk (i*gridDim + b) of elems_per_cta_iter elements
  const double2 v = make_double2(threadIdx.x, blockIdx.x);
  for (int i = 0; i < iters; ++i) {
    size_t base = (((size_t)i * gridDim.x + blockIdx.x) * elems_per_cta_iter) % total_elems;
    for (size_t e = threadIdx.x; e < elems_per_cta_iter; e += blockDim.x) dst[base + e] = v;
  }
}

// the same sweep with 32 bytes per lane (st.global.v4.f64, sm_100+): does a wider store per thread change the NVLink packets?
__global__ void k_stg32(double4 *dst, size_t elems_per_cta_iter, int iters, size_t total_elems) {
  const double a = threadIdx.x, b = blockIdx.x;
  for (int i = 0; i < iters; ++i) {
    size_t base = (((size_t)i * gridDim.x + blockIdx.x) * elems_per_cta_iter) % total_elems;
    for (size_t e = threadIdx.x; e < elems_per_cta_iter; e += blockDim.x)
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + base + e), "d"(a), "d"(b), "d"(a), "d"(b) : "memory");
  }
}

__global__ void k_bulk(char *dst, int chunk_bytes, int chunks_per_tile, int iters, size_t total_bytes, int max_groups) {
  extern __shared__ __align__(128) char sm[];
  for (int j = threadIdx.x; j < chunk_bytes * chunks_per_tile / 16; j += blockDim.x) ((double2 *)sm)[j] = make_double2(j, blockIdx.x);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const size_t tile_bytes = (size_t)chunk_bytes * chunks_per_tile;
  if (threadIdx.x < chunks_per_tile) {
    for (int i = 0; i < iters; ++i) {
      size_t base = (((size_t)i * gridDim.x + blockIdx.x) * tile_bytes) % total_bytes;
      char *g = dst + base + (size_t)threadIdx.x * chunk_bytes;
      unsigned s = (unsigned)__cvta_generic_to_shared(sm + (size_t)threadIdx.x * chunk_bytes);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (max_groups == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (max_groups == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t bytes = 1ull << 30;
  char *buf[2];
  cudaStream_t st[2];
  cudaEvent_t e0[2], e1[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&buf[d], bytes));
    CK(cudaStreamCreate(&st[d]));
    CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
  }
  CK(cudaSetDevice(0));
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaSetDevice(1));
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const size_t total = 4ull << 30;   // bytes written per direction per measurement
  auto run = [&](const char *name, int bidir, auto launch) {
    for (int rep = 0; rep < 2; ++rep) {
      for (int d = 0; d <= bidir; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(e0[d], st[d])); launch(d); CK(cudaEventRecord(e1[d], st[d])); }
      for (int d = 0; d <= bidir; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
    }
    float ms = 0, m;
    for (int d = 0; d <= bidir; ++d) { CK(cudaEventElapsedTime(&m, e0[d], e1[d])); ms = m > ms ? m : ms; }
    printf("%-44s %s  %7.3f ms  %7.1f GB/s per direction\n", name, bidir ? "bidir" : "unidir", ms, total / ms / 1e6);
  };
  for (int bidir = 0; bidir <= 1; ++bidir) {
    for (int ctas : {148, 296, 592}) {
      for (int threads : {256, 512}) {
        char name[128];
        const size_t chunk_elems = 8192 / 16;   // 8 KB per CTA iteration
        const int iters = (int)(total / (8192ull * ctas));
        snprintf(name, sizeof(name), "stg  %d CTAs x %d thr, 8 KB chunks", ctas, threads);
        run(name, bidir, [&](int d) { k_stg<<<ctas, threads, 0, st[d]>>>((double2 *)buf[1 - d], chunk_elems, iters, bytes / 16); });
      }
    }
    for (int ctas : {148, 296}) {
      for (int threads : {256, 512}) {
        char name[128];
        const size_t chunk_elems = 8192 / 32;   // 8 KB per CTA iteration, 32 B per lane
        const int iters = (int)(total / (8192ull * ctas));
        snprintf(name, sizeof(name), "stg32 %d CTAs x %d thr, 8 KB chunks, 32 B per lane", ctas, threads);
        run(name, bidir, [&](int d) { k_stg32<<<ctas, threads, 0, st[d]>>>((double4 *)buf[1 - d], chunk_elems, iters, bytes / 32); });
      }
    }
    for (int ctas : {148, 296}) {
      for (int chunk : {2048, 8192, 32768}) {
        for (int groups : {0, 1, 3}) {
          char name[128];
          const int cpt = 65536 / chunk > 32 ? 32 : 65536 / chunk;
          const int iters = (int)(total / ((size_t)chunk * cpt * ctas));
          snprintf(name, sizeof(name), "bulk %d CTAs, %d B x %d per tile, %d groups in flight", ctas, chunk, cpt, groups);
          run(name, bidir, [&](int d) { k_bulk<<<ctas, 64, (size_t)chunk * cpt, st[d]>>>(buf[1 - d], chunk, cpt, iters, bytes, groups); });
        }
      }
    }
  }
  // reference: copy engine
  CK(cudaSetDevice(0));
  CK(cudaEventRecord(e0[0], st[0]));
  for (int i = 0; i < 4; ++i) CK(cudaMemcpyPeerAsync(buf[1], 1, buf[0], 0, bytes, st[0]));
  CK(cudaEventRecord(e1[0], st[0]));
  CK(cudaStreamSynchronize(st[0]));
  float ms; CK(cudaEventElapsedTime(&ms, e0[0], e1[0]));
  printf("%-44s unidir  %7.3f ms  %7.1f GB/s\n", "cudaMemcpyPeerAsync 4 x 1 GiB", ms, 4.0 * bytes / ms / 1e6);
  return 0;
}
