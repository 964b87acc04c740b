// tools/svd_sweeps.cpp -- host build of the device Jacobi SVD (tests/cpp/cuda_host_shims.h): mean number of fp64 sweeps per
// matrix with and without the single-precision preconditioner.
//   g++ -O2 -std=c++17 -ffp-contract=off -Wno-unknown-pragmas -DPPX_SVD_COUNT_SWEEPS [-DPPX_SVD_PRECOND=0] tools/svd_sweeps.cpp
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../tests/cpp/cuda_host_shims.h"
#include "../matrix_b200/csrc/ppx_math.cuh"
using namespace ppx_dev;
int main(){
  std::mt19937_64 rng(1); std::uniform_real_distribution<double> U(-1,1);
  const int n=20000; g_ppx_svd_sweeps=0;
  double maxorth=0;
  for(int i=0;i<n;i++){ double a[36],v[36],w2[6]; for(double&x:a)x=U(rng); jacobi_svd_core<6,6,true>(a,v,w2);
     for(int p=0;p<6;p++)for(int q=p;q<6;q++){double d=0;for(int k=0;k<6;k++)d+=v[k+6*p]*v[k+6*q]; d-= (p==q); if(fabs(d)>maxorth)maxorth=fabs(d);} }
  printf("svd6x6 mean fp64 sweeps %.3f  max |VtV-I| %.2e\n",(double)g_ppx_svd_sweeps/n,maxorth);
  g_ppx_svd_sweeps=0;
  for(int i=0;i<n;i++){ double a[12],v[9],w2[3]; for(double&x:a)x=U(rng); jacobi_svd_core<4,3,true>(a,v,w2);} 
  printf("svd4x3 mean fp64 sweeps %.3f\n",(double)g_ppx_svd_sweeps/n);
  g_ppx_svd_sweeps=0;
  for(int i=0;i<n;i++){ double a[36],b[6],x[6]; for(double&t:a)t=U(rng); for(double&t:b)t=U(rng); jacobi_rowsolve<6>(a,b,x);} 
  printf("rowsolve6 mean fp64 sweeps %.3f\n",(double)g_ppx_svd_sweeps/n);
}
