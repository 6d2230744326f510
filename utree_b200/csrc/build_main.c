/* utree-build_gg: the BUILD_GG binary of the reference makefile (itree.c:1379-1408), on the GPU. */
#include "../../include/utree_b200.h"
int main(int argc, char **argv) { return utb_build_main(argc, argv); }
