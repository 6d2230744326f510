"""Bench / test tooling (synthetic genomes, CTR files, reads).  Not part of the search path."""
