// Probe of TMA (cp.async.bulk.tensor) for FP64 tiles on sm_100a: verifies the tensor-map encoding, the mbarrier
// completion protocol and prints where each element of a box lands in shared memory for every swizzle mode.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu   (driver entry point, no -lcuda)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_LOOP;\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// map passed by value (__grid_constant__) or by pointer to global memory (use_global)
__global__ void probe2d(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_global, int c0, int c1, int box_elems, double* out) {
  extern __shared__ __align__(1024) double tile[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, box_elems * 8);
    tma_load_2d(tile, use_global ? gmap : &pmap, c0, c1, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out[i] = tile[i];
}
__global__ void probe3d(const __grid_constant__ CUtensorMap pmap, int c0, int c1, int c2, int box_elems, double* out) {
  extern __shared__ __align__(1024) double tile[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, box_elems * 8);
    tma_load_3d(tile, &pmap, c0, c1, c2, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out[i] = tile[i];
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int R = 96, C = 80;  // rows, cols (cols contiguous)
  std::vector<double> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 1000 + c;
  double *d, *out;
  CK(cudaMalloc(&d, sizeof(double) * R * C));
  CK(cudaMemcpy(d, h.data(), sizeof(double) * R * C, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, sizeof(double) * 8192));
  CUtensorMap* gmap;
  CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
  const int BR = 16, BC = 16;  // box: 16 rows x 16 cols (128 bytes per row)
  const char* names[] = {"NONE", "32B", "64B", "128B", "128B_ATOM_32B", "128B_ATOM_32B_FLIP_8B", "128B_ATOM_64B"};
  for (int sw = 0; sw <= 5; ++sw) {
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 8};
    cuuint32_t box[2] = {(cuuint32_t)((sw == 1) ? 4 : (sw == 2) ? 8 : BC), BR};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== swizzle %s: encode rc=%d, box %u x %u\n", names[sw], (int)r, box[1], box[0]);
    if (r != CUDA_SUCCESS) continue;
    CK(cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice));
    int elems = box[0] * box[1];
    for (int ug = 0; ug < 2; ++ug) {
      CK(cudaMemset(out, 0, sizeof(double) * 8192));
      probe2d<<<1, 128, 16384>>>(map, gmap, ug, /*c0 (col)*/ 32, /*c1 (row)*/ 8, elems, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  kernel failed (use_global=%d): %s\n", ug, cudaGetErrorString(e)); return 1; }
      std::vector<double> o(elems);
      CK(cudaMemcpy(o.data(), out, sizeof(double) * elems, cudaMemcpyDeviceToHost));
      if (ug == 1) { printf("  map from global memory: first elem %.0f (expect 8032)\n", o[0]); continue; }
      // print, per smem row (box[0] elements), the source column offsets (c-32) of each slot; rows should be r-8
      for (int i = 0; i < (int)box[1]; ++i) {
        printf("  smem row %2d (src row %2d):", i, (int)(o[i * box[0]] / 1000) - 8);
        for (int j = 0; j < (int)box[0]; ++j) printf(" %2d", (int)o[i * box[0] + j] % 1000 - 32);
        printf("\n");
        if (i == 8 && sw != 4) { printf("  ...\n"); break; }
      }
    }
  }
  // OOB zero fill: box starting at row 90 (rows 96.. are out of range) and col 72 (cols 80.. out of range)
  {
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 8};
    cuuint32_t box[2] = {16, 16}, estr[2] = {1, 1};
    encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    probe2d<<<1, 128, 16384>>>(map, gmap, 0, 72, 90, 256, out);
    CK(cudaDeviceSynchronize());
    std::vector<double> o(256);
    CK(cudaMemcpy(o.data(), out, sizeof(double) * 256, cudaMemcpyDeviceToHost));
    printf("== OOB: [0][0]=%.0f (expect 90072) [0][7]=%.0f (90079) [0][8]=%.0f (0) [5][0]=%.0f (95072) [6][0]=%.0f (0)\n", o[0], o[7], o[8], o[5 * 16], o[6 * 16]);
  }
  // 3D map over a (K=8, S=2, Y=40) tensor viewed from the same buffer: dims (inner->outer) {Y, S, K}; box {16, 2, 4}
  {
    CUtensorMap map;
    cuuint64_t gdim[3] = {40, 2, 8};
    cuuint64_t gstr[2] = {40 * 8, 80 * 8};
    cuuint32_t box[3] = {16, 2, 4}, estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== 3D encode rc=%d\n", (int)r);
    probe3d<<<1, 128, 16384>>>(map, /*y*/ 16, /*s*/ 0, /*k*/ 2, 128, out);
    CK(cudaDeviceSynchronize());
    std::vector<double> o(128);
    CK(cudaMemcpy(o.data(), out, sizeof(double) * 128, cudaMemcpyDeviceToHost));
    // linear index in buffer = k*80 + s*40 + y -> value = (idx/80)*1000 + idx%80 with C=80 => r=k, c=s*40+y
    printf("   smem[k=0][s=0][0]=%.0f (expect 2016) [k=0][s=1][0]=%.0f (2056) [k=1][s=0][3]=%.0f (3019)\n", o[0], o[16], o[32 + 3]);
  }
  printf("done\n");
  return 0;
}
