#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
#include <time.h>
#include <immintrin.h>
static double now(){struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){
  size_t n = (size_t)1<<30; /* 1G elements: 4 GB in, 8 GB out */
  int nt = argc>1?atoi(argv[1]):omp_get_max_threads();
  int32_t* in = (int32_t*)aligned_alloc(64, n*4); int64_t* out=(int64_t*)aligned_alloc(64,n*8);
  #pragma omp parallel for num_threads(nt)
  for(size_t i=0;i<n;i++){in[i]=(int32_t)i; }
  #pragma omp parallel for num_threads(nt)
  for(size_t i=0;i<n;i++){out[i]=0;}
  for(int rep=0;rep<3;rep++){
    double t0=now();
    #pragma omp parallel for num_threads(nt) schedule(static)
    for(size_t i=0;i<n;i+=8){
      __m256i x=_mm256_load_si256((const __m256i*)(in+i));
      __m256i lo=_mm256_cvtepi32_epi64(_mm256_castsi256_si128(x));
      __m256i hi=_mm256_cvtepi32_epi64(_mm256_extracti128_si256(x,1));
      _mm256_stream_si256((__m256i*)(out+i),lo); _mm256_stream_si256((__m256i*)(out+i+4),hi);
    }
    double t=now()-t0;
    printf("threads %d: widen %.2f GB out in %.3f s = %.1f GB/s written (%.1f GB/s total traffic)\n",nt,n*8/1e9,t,n*8/1e9/t,n*12/1e9/t);
  }
  return (int)(out[12345]&1);
}
