/* compress_main.c -- `utree-compress`: the reference CLI contract (itree.c:1352-1355) on top of the C ABI. */
#include "../../include/utree_b200.h"
int main(int argc, char **argv) { return utb_compress_main(argc, argv); }
