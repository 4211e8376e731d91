set -x
# session 3, call 15: 16-warp versions of the short-rule hexahedron shape (bulk rule of configs 3 / 5); points in flight of the Nitsche moment-set loop
for lib in sbase sxs3 smw6; do echo "== $lib"
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 128 1 stiffness 8 8 7 | cut -c1-140
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 128 1 mass 8 8 7 | cut -c1-140
done
for lib in sbase nu1 nu4; do echo "== $lib"
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 nitsche 4 16 7 | cut -c1-140
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 nitsche 8 27 7 | cut -c1-140
done
