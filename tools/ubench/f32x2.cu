#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(unsigned long long r, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(r)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c){ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
template<int MODE> __global__ void k(float* out, int iters, float s){
  float a[8]; for(int i=0;i<8;i++) a[i]=threadIdx.x*0.001f+i;
  unsigned long long p[4]; for(int i=0;i<4;i++) p[i]=pk(a[2*i],a[2*i+1]);
  unsigned long long ss=pk(s,s), tt=pk(0.5f,0.25f);
  for(int it=0;it<iters;it++){
    if(MODE==0){
#pragma unroll
      for(int i=0;i<8;i++) a[i]=fmaf(a[i],s,0.5f);
    } else {
#pragma unroll
      for(int i=0;i<4;i++) p[i]=fma2(p[i],ss,tt);
    }
  }
  float r=0; if(MODE==0){for(int i=0;i<8;i++) r+=a[i];} else {for(int i=0;i<4;i++){float x,y; upk(p[i],x,y); r+=x+y;}}
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
int main(){
  float* out; cudaMalloc(&out, 148*8*256*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int mode=0;mode<2;mode++){
    for(int rep=0;rep<2;rep++){
      cudaEventRecord(e0);
      if(mode==0) k<0><<<148*8,256>>>(out,100000,0.999f); else k<1><<<148*8,256>>>(out,100000,0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      double fma = 148.0*8*256*100000.0*8;
      printf("mode %d: %.3f ms  %.2f TFMA/s  (%s)\n", mode, ms, fma/ms/1e9, mode? "fma.rn.f32x2":"fma.rn.f32");
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
