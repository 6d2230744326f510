/* utree-search: the non-GG search binary of the reference makefile (itree.c with -D SEARCH). */
#include "../../include/utree_b200.h"
int main(int argc, char **argv) { return utb_main_shallow(argc, argv); }
