set -x
./profiles/microbench/row_store_bench
