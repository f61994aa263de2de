set -x
cd $GRAFT_REPO_ROOT
run_ncu() {  # name regex cmd...
  name=$1; regex=$2; shift 2
  "$@" > gpurun_out/r02_plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$regex -s 8 -c 2 -o gpurun_out/r02_$name "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  tail -1 gpurun_out/r02_ncu_$name.log
  python tools/ncu_digest.py gpurun_out/r02_$name.ncu-rep gpurun_out/r02_ncu_full_$name.txt > /dev/null
  python tools/ncu_traffic.py gpurun_out/r02_$name.ncu-rep $name > /dev/null && cp profiles/traffic.json gpurun_out/traffic.json
  rm -f gpurun_out/r02_$name.ncu-rep
}
run_ncu pr_fused pr_fused python tools/prof_target.py pr 256 3 12
run_ncu ew_gv ew_kernel python tools/prof_target.py gv 256 3 12
run_ncu sp_gv stencil_tma python tools/prof_target.py gv 256 3 12
run_ncu ew_pipe_r ew_kernel python tools/prof_target.py pipe_pr 256 3 12
run_ncu sp_pipe_r stencil_tma python tools/prof_target.py pipe_pr 256 3 12
run_ncu sp_hs stencil_tma python tools/prof_target.py hs 256 3 12
run_ncu csr_sp_pr csr_stream python tools/prof_csr_target.py pr 12
run_ncu csr_sp_pipe_r csr_stream python tools/prof_csr_target.py pipe_pr 12
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ls -la gpurun_out | tail -30; du -sh gpurun_out
