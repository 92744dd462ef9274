#!/bin/bash
# compute-sanitizer over the hot path (run on the GPU box): memcheck, racecheck, synccheck on
#   * __graft_entry__.smoke()                      (cc_step_tpe_kernel<8,4>: TMA image ring, policy bitmap aliased onto it)
#   * the small-lattice kernel, int8 rows          (cc_step_tpe2_kernel<8,1>: per-thread images, one bulk copy per group)
#   * the lane-group kernel with 64 agents, int8   (cc_kernel<32,2,1,step>: shifted template copies, per-warp staging)
# Only this library's kernels are instrumented (--kernel-regex); logs go to gpurun_out/ and are summarised on stdout.
out=${1:-gpurun_out}
mkdir -p $out
run() {  # tag, tool, command...
  tag=$1; tool=$2; shift 2
  compute-sanitizer --tool $tool --kernel-regex kns=ccb --error-exitcode 9 --log-file $out/sanitizer_${tag}_${tool}.log "$@" > $out/sanitizer_${tag}_${tool}.out 2>&1
  echo "$tag $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitizer_${tag}_${tool}.log | tail -1)"
}
for tool in memcheck racecheck synccheck; do
  run smoke $tool python -c "import __graft_entry__ as g; g.smoke()"
  run tpe2_int8 $tool python profiles/run_steps.py int8 waiting 6 4099
  run tpe2_table_fused $tool python profiles/run_steps.py table greedy 2 4099 5
  run lanes64_int8 $tool python profiles/run_steps.py int8 random 3 1027 1 large
done
