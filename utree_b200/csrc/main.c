/* main.c -- `utree-search_gg` / `utree-searchGG`: the reference CLI contract
 * (itree.c:1357-1377) on top of the C ABI. */
#include "../../include/utree_b200.h"
int main(int argc, char **argv) { return utb_main(argc, argv); }
