// DMMA (FP64 tensor) probe for sm_100a: verifies the fragment layouts we assume for
// mma.sync.*.f64 shapes and measures their peak issue rate against plain DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void mma884(double &c0,double &c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4],const double (&a)[2],double b){
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4],const double (&a)[4],const double (&b)[2]){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4],const double (&a)[8],const double (&b)[4]){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}

// ---- layout checks: A is MxK row-major, B is KxN (given as B[k][n]), C = A*B
__global__ void chk884(const double*A,const double*B,double*C){
  int l=threadIdx.x,g=l>>2,t=l&3; double c0=0,c1=0;
  mma884(c0,c1,A[g*4+t],B[t*8+g]); C[g*8+2*t]=c0; C[g*8+2*t+1]=c1;
}
__global__ void chk1684(const double*A,const double*B,double*C){
  int l=threadIdx.x,g=l>>2,t=l&3; double c[4]={0,0,0,0}; double a[2]={A[g*4+t],A[(g+8)*4+t]};
  mma1684(c,a,B[t*8+g]); C[g*8+2*t]=c[0]; C[g*8+2*t+1]=c[1]; C[(g+8)*8+2*t]=c[2]; C[(g+8)*8+2*t+1]=c[3];
}
__global__ void chk1688(const double*A,const double*B,double*C){
  int l=threadIdx.x,g=l>>2,t=l&3; double c[4]={0,0,0,0};
  double a[4]={A[g*8+t],A[(g+8)*8+t],A[g*8+t+4],A[(g+8)*8+t+4]}; double b[2]={B[t*8+g],B[(t+4)*8+g]};
  mma1688(c,a,b); C[g*8+2*t]=c[0]; C[g*8+2*t+1]=c[1]; C[(g+8)*8+2*t]=c[2]; C[(g+8)*8+2*t+1]=c[3];
}
__global__ void chk16816(const double*A,const double*B,double*C){
  int l=threadIdx.x,g=l>>2,t=l&3; double c[4]={0,0,0,0}; double a[8],b[4];
  for(int i=0;i<8;i++) a[i]=A[(g+8*(i&1))*16+t+4*(i>>1)];
  for(int i=0;i<4;i++) b[i]=B[(t+4*i)*8+g];
  mma16816(c,a,b); C[g*8+2*t]=c[0]; C[g*8+2*t+1]=c[1]; C[(g+8)*8+2*t]=c[2]; C[(g+8)*8+2*t+1]=c[3];
}

template<int MODE> __global__ void __launch_bounds__(256) rate(double*out,int iters,double seed){
  // 8 independent accumulator tiles per warp to cover the pipe latency
  double acc[8][4]; for(int i=0;i<8;i++)for(int j=0;j<4;j++)acc[i][j]=seed*(i+j);
  double a[8],b[4]; for(int i=0;i<8;i++)a[i]=seed+threadIdx.x*1e-9*i; for(int i=0;i<4;i++)b[i]=seed*0.5+i*1e-9;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++){
      if(MODE==0) mma884(acc[i][0],acc[i][1],a[i&7],b[i&3]);
      if(MODE==1){ double aa[2]={a[i&7],a[(i+1)&7]}; mma1684(acc[i],aa,b[i&3]); }
      if(MODE==2){ double aa[4]={a[i&7],a[(i+1)&7],a[(i+2)&7],a[(i+3)&7]}; double bb[2]={b[i&3],b[(i+1)&3]}; mma1688(acc[i],aa,bb); }
      if(MODE==3){ mma16816(acc[i],a,b); }
      if(MODE==5){ // mixed: one m8n8k4 DMMA (512 flop/warp) + 2 DFMA per lane (128 flop/warp): do the pipes overlap?
        mma884(acc[i][0],acc[i][1],a[i&7],b[i&3]);
        acc[i][2]=fma(acc[i][2],a[2],b[2]); acc[i][3]=fma(acc[i][3],a[3],b[3]);
      }
      if(MODE==6){ // mixed 1:1 flops: one m8n8k4 DMMA (512) + 8 DFMA per lane (512)
        mma884(acc[i][0],acc[i][1],a[i&7],b[i&3]);
        #pragma unroll
        for(int j=0;j<4;j++){ acc[(i+1)&7][2]=fma(acc[(i+1)&7][2],a[j],b[j]); acc[(i+2)&7][3]=fma(acc[(i+2)&7][3],a[j+4],b[j]); }
      }
      if(MODE==4){ // plain DFMA: 32 independent chains
        #pragma unroll
        for(int j=0;j<4;j++) acc[i][j]=fma(acc[i][j],a[j],b[j]);
      }
    }
  }
  double s=0; for(int i=0;i<8;i++)for(int j=0;j<4;j++)s+=acc[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int MODE> double run_rate(const char*name,double flop_per_warp_inst,int ctas_per_sm){
  int dev; CK(cudaGetDevice(&dev)); cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int grid=p.multiProcessorCount*ctas_per_sm, iters=20000; double*out; CK(cudaMalloc(&out,sizeof(double)*grid*256));
  rate<MODE><<<grid,256>>>(out,100,1.0); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e30f;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); rate<MODE><<<grid,256>>>(out,iters,1.0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  double warps=(double)grid*8, insts=warps*iters*8.0; double tf=insts*flop_per_warp_inst/(best*1e-3)/1e12;
  printf("{\"probe\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n",name,ctas_per_sm,best,tf);
  cudaFree(out); return tf;
}

template<class K> double check(K kern,int M,int N,int Kd){
  std::vector<double> A(M*Kd),B(Kd*N),C(M*N),R(M*N,0.0);
  for(auto&x:A)x=rand()/(double)RAND_MAX-0.5; for(auto&x:B)x=rand()/(double)RAND_MAX-0.5;
  for(int i=0;i<M;i++)for(int j=0;j<N;j++){double s=0;for(int k=0;k<Kd;k++)s+=A[i*Kd+k]*B[k*N+j];R[i*N+j]=s;}
  double*dA,*dB,*dC; CK(cudaMalloc(&dA,A.size()*8));CK(cudaMalloc(&dB,B.size()*8));CK(cudaMalloc(&dC,C.size()*8));
  cudaMemcpy(dA,A.data(),A.size()*8,cudaMemcpyHostToDevice);cudaMemcpy(dB,B.data(),B.size()*8,cudaMemcpyHostToDevice);
  kern<<<1,32>>>(dA,dB,dC); CK(cudaDeviceSynchronize()); cudaMemcpy(C.data(),dC,C.size()*8,cudaMemcpyDeviceToHost);
  double e=0; for(size_t i=0;i<C.size();i++)e=fmax(e,fabs(C[i]-R[i])); cudaFree(dA);cudaFree(dB);cudaFree(dC); return e;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n",p.name,p.multiProcessorCount,p.clockRate);
  printf("{\"layout_err\":{\"m8n8k4\":%.2e,\"m16n8k4\":%.2e,\"m16n8k8\":%.2e,\"m16n8k16\":%.2e}}\n",
    check(chk884,8,8,4),check(chk1684,16,8,4),check(chk1688,16,8,8),check(chk16816,16,8,16));
  for(int c=1;c<=4;c*=2){
    run_rate<0>("dmma_m8n8k4",2.0*8*8*4,c); run_rate<1>("dmma_m16n8k4",2.0*16*8*4,c);
    run_rate<2>("dmma_m16n8k8",2.0*16*8*8,c); run_rate<3>("dmma_m16n8k16",2.0*16*8*16,c);
    run_rate<4>("dfma",2.0*32*4,c);
    run_rate<5>("mixed_dmma512_dfma128",512.0+128.0,c); run_rate<6>("mixed_dmma512_dfma512",512.0+512.0,c);
  }
  return 0;
}
