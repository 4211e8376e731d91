set -x
for lib in build/variants/libcq_l10.so build/variants/libcq_l8.so build/variants/libcq_l10n.so; do echo "== $lib"
for r in "100 120" "40 60" "20 30" "8 27"; do set -- $r; CQ_LIB=$lib python profiles/one_form.py hex 96 1 stiffness $1 $2 5 | cut -c1-140; done
CQ_LIB=$lib python profiles/one_form.py hex 96 1 nitsche 20 30 5 | cut -c1-140
CQ_LIB=$lib python profiles/one_form.py hex 96 1 nitsche 8 27 5 | cut -c1-140
done
CQ_LIB=build/variants/libcq_l10.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hexahedra" 2>&1 | tail -2
