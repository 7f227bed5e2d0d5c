// Issue-rate probe for sm_100a: cycles per warp-instruction per SMSP for a few instruction classes, with W warps
// per SMSP and 8 independent register chains per thread.  Build: nvcc -arch=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITERS 4096
template <int OP>
__global__ void probe(float *out, long long *cyc, float a, float b, int ib) {
  float x[8]; int xi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = a + i + threadIdx.x; xi[i] = ib + i + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < N_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = x[i] + b;                                  // FADD
      if (OP == 1) x[i] = fmaf(x[i], a, b);                          // FFMA
      if (OP == 2) x[i] = __saturatef(x[i] - b);                     // FADD.SAT
      if (OP == 3) xi[i] = (xi[i] ^ ib) + 0x1234;                    // LOP3 + IADD (ALU)
      if (OP == 4) { x[i] = x[i] + b; xi[i] = xi[i] ^ (ib + i); }    // FADD + LOP3 (both pipes)
      if (OP == 5) { if (xi[i] > ib) x[i] += b; xi[i] += 3; }        // ISETP + @P FADD + IADD
      if (OP == 6) x[i] = fmaf(x[i], 2.0f, b);                       // FFMA with immediate
    }
  }
  long long t1 = clock64();
  float s = 0; int si = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += x[i]; si += xi[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + si;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  const char *names[] = {"FADD", "FFMA", "FADD.SAT", "LOP3+IADD", "FADD+LOP3", "ISETP+@FADD+IADD", "FFMA imm"};
  int per_iter[] = {8, 8, 8, 16, 16, 24, 8};
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    int threads = warps_per_smsp * 4 * 32;
    for (int op = 0; op < 7; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: probe<0><<<148, threads>>>(out, cyc, 1.f, 1e-3f, 7); break;
          case 1: probe<1><<<148, threads>>>(out, cyc, 1.0001f, 1e-3f, 7); break;
          case 2: probe<2><<<148, threads>>>(out, cyc, 1.f, 1e-3f, 7); break;
          case 3: probe<3><<<148, threads>>>(out, cyc, 1.f, 1e-3f, 7); break;
          case 4: probe<4><<<148, threads>>>(out, cyc, 1.f, 1e-3f, 7); break;
          case 5: probe<5><<<148, threads>>>(out, cyc, 1.f, 1e-3f, 7); break;
          case 6: probe<6><<<148, threads>>>(out, cyc, 1.f, 1e-3f, 7); break;
        }
        cudaDeviceSynchronize();
      }
      double instr_per_smsp = (double)N_ITERS * per_iter[op] * warps_per_smsp;
      printf("warps/SMSP %d  %-18s cycles/warp-instr/SMSP = %.3f\n", warps_per_smsp, names[op], *cyc / instr_per_smsp);
    }
  }
  return 0;
}
