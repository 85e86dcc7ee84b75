python -m pytest tests -m gpu -x -q 2>&1 | tail -3
scripts/ab_variants.sh gpurun_out/r02_ab_specfrs.txt 2 build/variants/lib_head.so ratilqr.jl_b200/csrc/libratilqr_b200.so
