// microbench_pipes.cu -- measured per-SM instruction throughput of the pipes K1 leans on (B200).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench_pipes tools/microbench_pipes.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int OP>
__global__ void k_fp64(double* out, int iters, double a0) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
  const double b = 1.000000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma(a[i], b, c);
      if (OP == 1) a[i] = a[i] + c;
      if (OP == 2) a[i] = a[i] * b;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_f2f(double* out, int iters, float a0) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
  double s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double d = (double)a[i];          // F2F.F64.F32
      a[i] = (float)(d) + 1.0f;         // F2F.F32.F64 + FADD
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds128(double* out, int iters) {
  __shared__ double2 buf[2304];
  for (int i = threadIdx.x; i < 2304; i += blockDim.x) buf[i] = make_double2(i, -i);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double2 v = buf[(idx + 144 * i) % 2304];
      acc.x += v.x;
      acc.y += v.y;
      buf[(idx + 144 * i + 7) % 2304] = acc;  // STS.128
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* out;
  cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 4);
  const int iters = 20000;
  const char* names[3] = {"DFMA", "DADD", "DMUL"};
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int threads = 32 * warps, blocks = sms;  // one CTA per SM
    float ms[3];
    ms[0] = time_ms([&] { k_fp64<0><<<blocks, threads>>>(out, iters, 1.0); });
    ms[1] = time_ms([&] { k_fp64<1><<<blocks, threads>>>(out, iters, 1.0); });
    ms[2] = time_ms([&] { k_fp64<2><<<blocks, threads>>>(out, iters, 1.0); });
    for (int o = 0; o < 3; ++o) {
      double inst = (double)iters * 8 * threads;  // thread-instructions per SM
      double clk = ms[o] * 1e-3 * khz * 1e3;
      printf("%s  warps/SM %2d : %.1f thread-instr / clk / SM (at nominal %d MHz)\n", names[o], warps, inst / clk, khz / 1000);
    }
    float mf = time_ms([&] { k_f2f<<<blocks, threads>>>(out, iters, 1.0f); });
    printf("F2F pair (f32->f64->f32) warps/SM %2d : %.1f conversions / clk / SM\n", warps, (double)iters * 16 * threads / (mf * 1e-3 * khz * 1e3));
    float ml = time_ms([&] { k_lds128<<<blocks, threads>>>(out, iters / 10); });
    printf("LDS.128+STS.128 warps/SM %2d : %.1f bytes / clk / SM\n", warps, (double)(iters / 10) * 8 * 32.0 * threads / (ml * 1e-3 * khz * 1e3));
  }
  return 0;
}
