#include <math.h>
#include <stdint.h>
/* out[m,p] = sum_k a[m,k] b[k,p] with strides, sequential fma chain */
void seq_fma_mm(const double*a,int64_t sam,int64_t sak,const double*b,int64_t sbk,int64_t sbp,double*out,int64_t M,int64_t K,int64_t P){
  for(int64_t m=0;m<M;m++)for(int64_t p=0;p<P;p++){double acc=0.0;for(int64_t k=0;k<K;k++)acc=fma(a[m*sam+k*sak],b[k*sbk+p*sbp],acc);out[m*P+p]=acc;}
}
void seq_mm(const double*a,int64_t sam,int64_t sak,const double*b,int64_t sbk,int64_t sbp,double*out,int64_t M,int64_t K,int64_t P){
  for(int64_t m=0;m<M;m++)for(int64_t p=0;p<P;p++){double acc=0.0;for(int64_t k=0;k<K;k++)acc=acc+a[m*sam+k*sak]*b[k*sbk+p*sbp];out[m*P+p]=acc;}
}
