export DFF_B200_LIB=tools/libdff_trace.so DFF_SLAB_TRACE=1
for args in "8 8 1 3 3 1 4 5 256 256" "16 8 3 3 3 1 4 5 256 256" "64 64 3 3 3 1 4 5 32 32" "32 32 3 3 3 1 4 5 64 64" "128 128 3 3 3 1 4 5 8 8"; do
  echo "=== $args"; timeout 120 python tools/trace_conv.py $args 2>&1 | grep -E "^trace|CTA 0" | tail -2
done
