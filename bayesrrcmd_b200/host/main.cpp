// brr_gpu: the reference's `brr` command line for --analysis-type ppbayes / asyncppbayes on the B200 engine
// (reference src/main.cpp).  Example:
//   brr_gpu --data-file test/data/uk10k_chr1_1mb.bed --pheno test/data/test.phen --analysis-type ppbayes \
//           --chain-length 10 --burn-in 0 --thin 1 --mcmc-samples out.csv --S 0.01,0.001,0.0001 --seed 1
#include "brr_host.hpp"

int main(int argc, const char *argv[])
{
    return brr_host_main(argc, argv);
}
