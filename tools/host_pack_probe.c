// host scan/pack feasibility: T threads, each OR-reduces + packs (2 bits per u16) a slice of a 33.5 MB buffer
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#define NCH 4096
static uint16_t* buf; static uint8_t* out; static int T;
static pthread_barrier_t bar;
static double now(){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
__attribute__((target("avx2"))) static uint32_t pack_chunk(const uint16_t* v, uint8_t* o){
  __m256i acc=_mm256_setzero_si256();
  for(int i=0;i<4096;i+=64){
    __m256i a=_mm256_loadu_si256((const __m256i*)(v+i)),b=_mm256_loadu_si256((const __m256i*)(v+i+16)),c=_mm256_loadu_si256((const __m256i*)(v+i+32)),d=_mm256_loadu_si256((const __m256i*)(v+i+48));
    acc=_mm256_or_si256(acc,_mm256_or_si256(_mm256_or_si256(a,b),_mm256_or_si256(c,d)));
    // 2-bit pack: a | b<<2 | c<<4 | d<<6 per 16-bit lane -> low byte holds 4 voxels (voxel order interleaved by 16; the device undoes it)
    __m256i p=_mm256_or_si256(_mm256_or_si256(a,_mm256_slli_epi16(b,2)),_mm256_or_si256(_mm256_slli_epi16(c,4),_mm256_slli_epi16(d,6)));
    __m256i q=_mm256_packus_epi16(p,p); // low bytes
    _mm_storel_epi64((__m128i*)(o+(i>>2)),_mm256_castsi256_si128(q));
    _mm_storel_epi64((__m128i*)(o+(i>>2)+8),_mm256_extracti128_si256(q,1));
  }
  __m128i r=_mm_or_si128(_mm256_castsi256_si128(acc),_mm256_extracti128_si256(acc,1));
  r=_mm_or_si128(r,_mm_srli_si128(r,8)); r=_mm_or_si128(r,_mm_srli_si128(r,4)); r=_mm_or_si128(r,_mm_srli_si128(r,2));
  return _mm_cvtsi128_si32(r)&0xFFFF;
}
static void* work(void* arg){
  long t=(long)arg; volatile uint32_t sink=0;
  for(int rep=0;rep<22;++rep){
    pthread_barrier_wait(&bar);
    for(int c=t;c<NCH;c+=T) sink|=pack_chunk(buf+(size_t)c*4096,out+(size_t)c*1024);
    pthread_barrier_wait(&bar);
  }
  return 0;
}
int main(int argc,char**argv){
  T=argc>1?atoi(argv[1]):16;
  buf=aligned_alloc(4096,(size_t)NCH*8192); out=aligned_alloc(4096,(size_t)NCH*1024);
  for(size_t i=0;i<(size_t)NCH*4096;++i) buf[i]=(i*2654435761u>>13)&3;
  pthread_barrier_init(&bar,0,T+1);
  pthread_t th[64]; for(long t=0;t<T;++t) pthread_create(&th[t],0,work,(void*)t);
  double best=1e9;
  for(int rep=0;rep<22;++rep){ pthread_barrier_wait(&bar); double t0=now(); pthread_barrier_wait(&bar); double dt=now()-t0; if(rep>1&&dt<best)best=dt; }
  for(int t=0;t<T;++t) pthread_join(th[t],0);
  printf("threads %d: best %.1f us for %d chunks (%.1f GB/s read)\n",T,best*1e6,NCH,NCH*8192.0/best/1e9);
  return 0;
}
