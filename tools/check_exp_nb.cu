#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "../bayesrrcmd_b200/csrc/sampler.cuh"
int main(){ double worst=0; srand(1); for(int i=0;i<4000000;i++){ double x=(rand()/(double)RAND_MAX-0.5)*1400.0; if(i%4==0) x=(rand()/(double)RAND_MAX-0.5)*2.0; if (i%7==0) x = -rand()/(double)RAND_MAX*40; double a=brr::exp_nb(x), b=exp(x); double e=fabs(a-b)/b; if(e>worst){worst=e;} } printf("worst rel err %.3e  exp_nb(0)=%.17g exp_nb(-700)=%.6e vs %.6e  exp_nb(700)=%.6e vs %.6e\n", worst, brr::exp_nb(0.0), brr::exp_nb(-700.0), exp(-700.0), brr::exp_nb(700.0), exp(700.0)); }
