#!/bin/bash
# A/B of experimental builds of libkpapa.so: tools/ab_libs.sh _suffix1 _suffix2 ...   ("" = the in-tree build)
# (build a variant with nvcc ... -o kmerpapa_b200/libkpapa_suffix.so; KP_LIBKPAPA selects it)
for rep in 1 2; do
for v in "" "$@"; do
  if [ -z "$v" ]; then unset KP_LIBKPAPA; else export KP_LIBKPAPA=$PWD/kmerpapa_b200/libkpapa$v.so; fi
  echo "variant [$v] $(timeout 120 python tools/profile_dp.py single NNNNANNNN 6 2>&1 | tail -3 | awk '{print $5}' | tr '\n' ' ')"
done; done
