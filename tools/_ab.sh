#!/bin/bash
# Developer A/B helper (not a test): time several builds of the library on the same GPU.
#   tests/_ab.sh torchpiv_b200/lib_A.so torchpiv_b200/lib_B.so ...
for lib in "$@"; do
  echo "== $lib"
  PIVB200_LIB=$PWD/$lib python tools/_sweep.py 16 2>&1 | tail -1
done
