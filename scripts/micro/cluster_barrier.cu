// micro-benchmark: what one level barrier of K3 costs when a tile is spread over a cluster of CTAs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_barrier cluster_barrier.cu && ./cluster_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE>
__global__ void __launch_bounds__(256) bar_kernel(unsigned *flags, double2 *boxes, long long *out, int iters)
{
  cg::cluster_group cl = cg::this_cluster();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE >= 2) {                                     // what a row leaves behind: a store and a flag atomic
      const int k = (blockIdx.x * 256 + threadIdx.x + i * 7919) & 0xfffff;
      if ((threadIdx.x & 31) == 0) { boxes[k] = make_double2(1.0, 2.0); atomicOr(flags + k, 1u << (i & 31)); }
    }
    if (MODE >= 1) asm volatile("fence.proxy.async;" ::: "memory");
    if (MODE == 3) __syncthreads(); else cl.sync();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE> void run(int cluster, const char *what, unsigned *flags, double2 *boxes, long long *out)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(32 * cluster); cfg.blockDim = dim3(256);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, bar_kernel<MODE>, flags, boxes, out, iters);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("cluster %d  %-52s %.2f us per barrier (%s)\n", cluster, what, h / 1965.0 / iters, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  unsigned *flags; double2 *boxes; long long *out;
  cudaMalloc(&flags, 4 << 20); cudaMalloc(&boxes, 16 << 20); cudaMalloc(&out, 8);
  cudaMemset(flags, 0, 4 << 20);
  for (int c : {1, 2, 4, 8}) {
    run<0>(c, "cluster.sync alone", flags, boxes, out);
    run<1>(c, "fence.proxy.async + cluster.sync", flags, boxes, out);
    run<2>(c, "store + atomicOr + fence + cluster.sync", flags, boxes, out);
  }
  run<3>(1, "store + atomicOr + fence + __syncthreads", flags, boxes, out);
  return 0;
}
