set -x
# session 3, call 18: moment warps of the medium hexahedron shape (means of 13..32 points per cell): 4 + {4, 6, 8} + 4
for lib in mbase mid6 mid8; do echo "== $lib"
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 stiffness 8 27 7 | cut -c1-140
CQ_LIB=build/variants/libcq_$lib.so python profiles/one_form.py hex 96 1 stiffness 20 30 7 | cut -c1-140
done
