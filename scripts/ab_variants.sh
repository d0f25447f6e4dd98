run() { timeout 300 python bench.py --no-cpu-baseline --steps 30 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stages_ms'].items()}, d['pilot']['replay_exact'])"; }
run base
touch dlsc_gc_planner_b200/csrc/dlsc_kernels_exact.cu; make -C dlsc_gc_planner_b200/csrc EXTRA=-DDLSC_SFC_MINB=6 > /dev/null 2>&1; run sfc_minb6
touch dlsc_gc_planner_b200/csrc/dlsc_kernels_exact.cu; make -C dlsc_gc_planner_b200/csrc EXTRA=-DDLSC_SFC_MINB=4 > /dev/null 2>&1; run sfc_minb4
touch dlsc_gc_planner_b200/csrc/dlsc_kernels_exact.cu dlsc_gc_planner_b200/csrc/dlsc_kernels_qp.cu; make -C dlsc_gc_planner_b200/csrc EXTRA="-DDLSC_GI_THREADS=128" > /dev/null 2>&1; run gi128
touch dlsc_gc_planner_b200/csrc/dlsc_kernels_exact.cu dlsc_gc_planner_b200/csrc/dlsc_kernels_qp.cu; make -C dlsc_gc_planner_b200/csrc EXTRA="-DDLSC_GI_THREADS=32" > /dev/null 2>&1; run gi32
