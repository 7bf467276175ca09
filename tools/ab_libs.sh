for rep in 1 2; do
for v in "" _w12 _w10 _w8; do
  if [ -z "$v" ]; then unset KP_LIBKPAPA; else export KP_LIBKPAPA=$PWD/kmerpapa_b200/libkpapa$v.so; fi
  echo "variant [$v] $(timeout 120 python tools/profile_dp.py single NNNNANNNN 6 2>&1 | tail -3 | awk '{print $5}' | tr '\n' ' ')"
done; done
