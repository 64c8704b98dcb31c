#!/bin/bash
# ncu --set full captures of the kernels BASELINE's north_star asks about (judge item: tensor-pipe utilisation of
# the GEMM phases, achieved HBM GB/s of rounding / scale search) -> gpurun_out/r2_ncu_<case>.csv (raw page, CSV).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export SLK_SWEEP_COMPACT=1
cap() { # case, kernel regex, launch-skip, launch-count
  timeout 600 ncu --set full --clock-control none -k "regex:$2" --launch-skip $3 --launch-count $4 --csv --page raw \
    --log-file gpurun_out/r2_ncu_$1.csv python tools/prof_kernels.py $1 2 > gpurun_out/r2_ncu_$1.log 2>&1
}
cap gemm4096 "tc_gemm_kernel" 1 1
cap c5gemm "tc_gemm_kernel" 1 1
cap k1 "tc_gemm_kernel" 1 1
cap k6 "tc_gemm_kernel" 1 1
cap k4 "round_f32" 2 2
cap search "scale_search_tab" 1 1
cap chol12 "chol_dag_kernel" 1 1
cap chol1 "chol_dag_kernel" 2 2
cap sweep "sweep_macro_kernel" 12 2
cap fullh "tc_gemm_kernel" 2 1
ls -la gpurun_out/r2_ncu_*.csv
# source-level (SASS) stall sampling of the macro sweep kernel and the batched Cholesky kernel
src() { # case, kernel regex, launch-skip
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$2" --launch-skip $3 --launch-count 1 \
    -f -o gpurun_out/r2_src_$1 python tools/prof_kernels.py $1 2 > gpurun_out/r2_src_$1.log 2>&1
  ncu -i gpurun_out/r2_src_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_src_$1_sass.csv 2>gpurun_out/r2_src_$1_sass.err
  ls -la gpurun_out/r2_src_$1.ncu-rep
}
src sweep "sweep_macro_kernel" 12
src chol12 "chol_dag_kernel" 1
du -sh gpurun_out
