// Microbenchmark: does packing the bicubic tap products as mul.rn.f32x2 (SASS FMUL2) buy issue slots
// on sm_100a, with the adds kept scalar so that ptxas cannot contract them into FFMA2?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2_probe f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mul2(float a, float b, float w, float& pa, float& pb) {
  unsigned long long x, y, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(w));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(pa), "=f"(pb) : "l"(r));
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(const float4* __restrict__ tex, const float* __restrict__ wts, float* out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  float ax = 0.f, ay = 0.f, az = 0.f;
  float4 tx[16];
  for (int i = 0; i < 16; ++i) tx[i] = tex[(t + i * 37) & 1023];
  float w[16];
  for (int i = 0; i < 16; ++i) w[i] = wts[(t + i) & 255];
  for (int it = 0; it < iters; ++it) {
    float rx = 0, ry = 0, rz = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float px, py, pz;
      if (MODE == 0) { px = tx[i].x * w[i]; py = tx[i].y * w[i]; pz = tx[i].z * w[i]; }
      else { mul2(tx[i].x, tx[i].y, w[i], px, py); pz = tx[i].z * w[i]; }
      if (i == 0) { rx = px; ry = py; rz = pz; } else { rx = rx + px; ry = ry + py; rz = rz + pz; }
    }
    ax += rx; ay += ry; az += rz;
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = w[i] + 1e-7f;   // keep the loop from being hoisted
  }
  out[t] = ax + ay + az;
}

int main() {
  float4* tex; float* w; float* out;
  cudaMalloc(&tex, 1024 * sizeof(float4)); cudaMalloc(&w, 256 * 4); cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMemset(tex, 0x3c, 1024 * sizeof(float4)); cudaMemset(w, 0x3c, 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) probe<0><<<148 * 8, 256>>>(tex, w, out, iters); else probe<1><<<148 * 8, 256>>>(tex, w, out, iters);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("mode %d (%s): %.3f ms\n", mode, mode ? "FMUL2 products + scalar adds" : "scalar", ms);
    }
  }
  float h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("%g\n", h[0]);
  return 0;
}
