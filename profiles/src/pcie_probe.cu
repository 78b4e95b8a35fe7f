// Probe: how fast can SMs read mapped pinned host memory over PCIe, by access pattern?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/bin/pcie_probe profiles/src/pcie_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void ldg_kernel(const uint4* __restrict__ src, int64_t nvec, unsigned* sink) {
  unsigned acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// segments of seg bytes every stride bytes (a footprint's rows), 16-byte loads, warp per segment run
__global__ void ldg_seg_kernel(const uint8_t* __restrict__ src, int64_t nseg, int seg, int stride, unsigned* sink) {
  unsigned acc = 0;
  const int vps = seg / 16;
  const int64_t total = nseg * vps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / vps; const int v = (int)(i - s * vps);
    uint4 x;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(src + s * stride + 16 * v));
    acc ^= x.x ^ x.y ^ x.z ^ x.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// the same segments fetched with 16-byte cp.async (LDGSTS) into shared memory, a footprint per warp
__global__ void ldgsts_seg_kernel(const uint8_t* __restrict__ src, int64_t nitems, int seg, int stride, int nrow, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint8_t* mine = smem + (size_t)wid * seg * nrow;
  const int vps = seg / 16, total = vps * nrow;
  unsigned acc = 0;
  for (int64_t it = (int64_t)blockIdx.x * nw + wid; it < nitems; it += (int64_t)gridDim.x * nw) {
    const uint8_t* p = src + it * (int64_t)stride * nrow;
    for (int idx = lane; idx < total; idx += 32) {
      const int r = idx / vps, v = idx - r * vps;
      const unsigned d = (unsigned)__cvta_generic_to_shared(mine + 16 * idx);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p + (int64_t)r * stride + 16 * v) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    acc ^= *reinterpret_cast<volatile unsigned*>(mine + 16 * lane);
  }
  if (acc == 0x12345678u) *sink = acc;
}

// coalesced 128-bit stores to mapped host memory (posted PCIe writes)
__global__ void stg_kernel(uint4* __restrict__ dst, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = make_uint4((unsigned)i, 1u, 2u, 3u);
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}
// one kernel reading rd_vec vectors from host and writing wr_vec vectors to host, interleaved by CTA parity
__global__ void duplex_kernel(const uint4* __restrict__ src, int64_t rd_vec, uint4* __restrict__ dst, int64_t wr_vec, unsigned* sink) {
  const int half = gridDim.x / 2;
  if (blockIdx.x & 1) {
    const int b = blockIdx.x / 2;
    for (int64_t i = (int64_t)b * blockDim.x + threadIdx.x; i < wr_vec; i += (int64_t)half * blockDim.x) {
      asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"((unsigned)i), "r"(1u), "r"(2u), "r"(3u) : "memory");
    }
  } else {
    const int b = blockIdx.x / 2;
    unsigned acc = 0;
    for (int64_t i = (int64_t)b * blockDim.x + threadIdx.x; i < rd_vec; i += (int64_t)half * blockDim.x) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
  }
}

__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned phase) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}

// one thread per CTA drives a ring of STAGES bulk copies; each "item" = nrow copies of seg bytes at stride
template <int STAGES>
__global__ void bulk_kernel(const uint8_t* __restrict__ src, int64_t nitems, int seg, int stride, int nrow, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[STAGES];
  const int item_bytes = seg * nrow;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned acc = 0;
    int64_t it = blockIdx.x;
    auto issue = [&](int64_t k, int s) {
      if (k < nitems) {
        mbar_expect(&bar[s], (unsigned)item_bytes);
        const uint8_t* p = src + k * (int64_t)stride * nrow;
        for (int r = 0; r < nrow; ++r) bulk_g2s(smem + (size_t)s * item_bytes + (size_t)r * seg, p + (int64_t)r * stride, (unsigned)seg, &bar[s]);
      }
    };
    for (int s = 0; s < STAGES; ++s) issue(it + (int64_t)s * gridDim.x, s);
    unsigned phase = 0; int s = 0;
    for (; it < nitems; it += gridDim.x) {
      mbar_wait(&bar[s], phase);
      acc ^= *reinterpret_cast<volatile unsigned*>(smem + (size_t)s * item_bytes);
      issue(it + (int64_t)STAGES * gridDim.x, s);
      if (++s == STAGES) { s = 0; phase ^= 1; }
    }
    if (acc == 0x12345678u) *sink = acc;
  }
}

int main() {
  const size_t bytes = 1ull << 30;
  uint8_t* h; CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  for (size_t i = 0; i < bytes; i += 4096) h[i] = (uint8_t)i;
  uint8_t* d; CK(cudaMalloc(&d, bytes));
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  auto report = [&](const char* name, double moved) {
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-58s %8.3f ms  %7.2f GB/s\n", name, ms, moved / ms / 1e6); fflush(stdout);
  };
  CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
  CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); report("copy engine H2D 1 GiB", (double)bytes);
  for (int tpb : {128, 256, 1024}) for (int cps : {1, 2, 4, 8}) {
    char nm[128]; snprintf(nm, 128, "LDG.128 coalesced, %d thr x %d CTA/SM", tpb, cps);
    ldg_kernel<<<148 * cps, tpb>>>((const uint4*)h, bytes / 16, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); ldg_kernel<<<148 * cps, tpb>>>((const uint4*)h, bytes / 16, sink); report(nm, (double)bytes);
  }
  for (int seg : {96, 192, 288, 384}) {
    const int stride = 672; const int64_t nseg = (int64_t)(bytes / stride) - 1;
    char nm[128]; snprintf(nm, 128, "LDG.128 segments %d B / stride %d, 256 thr x 4 CTA/SM", seg, stride);
    ldg_seg_kernel<<<148 * 4, 256>>>(h, nseg, seg, stride, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); ldg_seg_kernel<<<148 * 4, 256>>>(h, nseg, seg, stride, sink); report(nm, (double)nseg * seg);
  }
  for (int nw : {2, 6, 8}) {
    const int seg = 288, stride = 672, nrow = 64; const int64_t nitems = (int64_t)(bytes / ((int64_t)stride * nrow)) - 1;
    char nm[128]; snprintf(nm, 128, "LDGSTS.128 rows %d B / stride %d x %d rows, %d warps/SM", seg, stride, nrow, nw);
    CK(cudaFuncSetAttribute(ldgsts_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ldgsts_seg_kernel<<<148, nw * 32, nw * seg * nrow>>>(h, nitems, seg, stride, nrow, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); ldgsts_seg_kernel<<<148, nw * 32, nw * seg * nrow>>>(h, nitems, seg, stride, nrow, sink); report(nm, (double)nitems * seg * nrow);
  }
  // bulk copies: contiguous chunks
  for (int chunk : {512, 2048, 8192, 32768}) {
    const int64_t nitems = bytes / chunk;
    char nm[128]; snprintf(nm, 128, "cp.async.bulk contiguous %d B x 4 stages, 1 CTA/SM", chunk);
    CK(cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    bulk_kernel<4><<<148, 32, 4 * chunk>>>(h, nitems, chunk, chunk, 1, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); bulk_kernel<4><<<148, 32, 4 * chunk>>>(h, nitems, chunk, chunk, 1, sink); report(nm, (double)bytes);
  }
  for (int cps : {2, 4}) {
    const int chunk = 8192; const int64_t nitems = bytes / chunk;
    char nm[128]; snprintf(nm, 128, "cp.async.bulk contiguous %d B x 4 stages, %d CTA/SM", chunk, cps);
    CK(cudaEventRecord(a)); bulk_kernel<4><<<148 * cps, 32, 4 * chunk>>>(h, nitems, chunk, chunk, 1, sink); report(nm, (double)bytes);
  }
  // bulk copies: footprint-like, 75 rows of seg bytes at stride 672 per item
  for (int seg : {288, 240, 192}) for (int cps : {1, 2}) {
    const int stride = 672, nrow = 75; const int64_t nitems = (int64_t)(bytes / ((int64_t)stride * nrow)) - 1;
    char nm[128]; snprintf(nm, 128, "cp.async.bulk rows %d B / stride %d x 75 rows, 4 stages, %d CTA/SM", seg, stride, cps);
    CK(cudaFuncSetAttribute(bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    bulk_kernel<4><<<148 * cps, 32, 4 * seg * nrow>>>(h, nitems, seg, stride, nrow, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); bulk_kernel<4><<<148 * cps, 32, 4 * seg * nrow>>>(h, nitems, seg, stride, nrow, sink); report(nm, (double)nitems * seg * nrow);
  }
  // ---- the other direction and both at once
  uint8_t* h2; CK(cudaHostAlloc(&h2, bytes, cudaHostAllocDefault));
  cudaStream_t s2, s3; CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
  CK(cudaMemcpy(h2, d, bytes, cudaMemcpyDeviceToHost));
  CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost)); report("copy engine D2H 1 GiB", (double)bytes);
  for (int cps : {1, 4}) {
    char nm[128]; snprintf(nm, 128, "STG.128 coalesced to host, 256 thr x %d CTA/SM", cps);
    stg_kernel<<<148 * cps, 256>>>((uint4*)h2, bytes / 16); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); stg_kernel<<<148 * cps, 256>>>((uint4*)h2, bytes / 16); report(nm, (double)bytes);
  }
  {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    CK(cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost, s2));
    CK(cudaMemcpyAsync(d, h, bytes / 2, cudaMemcpyHostToDevice, s3));
    CK(cudaStreamSynchronize(s2)); CK(cudaStreamSynchronize(s3));
    report("copy engines: D2H 1 GiB || H2D 0.5 GiB (bytes = 1.5 GiB)", 1.5 * bytes);
  }
  {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    CK(cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost, s2));
    ldg_kernel<<<148 * 2, 256, 0, s3>>>((const uint4*)h, bytes / 32, sink);
    CK(cudaStreamSynchronize(s2)); CK(cudaStreamSynchronize(s3));
    report("copy engine D2H 1 GiB || LDG.128 zero-copy read 0.5 GiB", 1.5 * bytes);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    CK(cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost, s2));
    ldg_seg_kernel<<<148 * 4, 256, 0, s3>>>(h, (int64_t)(bytes / 672) - 1, 288, 672, sink);
    CK(cudaStreamSynchronize(s2)); CK(cudaStreamSynchronize(s3));
    report("copy engine D2H 1 GiB || LDG.128 288 B segments (0.43 GiB)", bytes + ((double)(bytes / 672) - 1) * 288);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); ldg_kernel<<<148 * 2, 256>>>((const uint4*)h, bytes / 32, sink); report("  (LDG.128 zero-copy read 0.5 GiB alone)", 0.5 * bytes);
  }
  for (int cps : {2, 4}) {
    char nm[128]; snprintf(nm, 128, "one kernel: LDG 0.5 GiB from host + STG 1 GiB to host, %d CTA/SM", cps);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); duplex_kernel<<<148 * cps, 256>>>((const uint4*)h, bytes / 32, (uint4*)h2, bytes / 16, sink); report(nm, 1.5 * bytes);
  }
  return 0;
}
