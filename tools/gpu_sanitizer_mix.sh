# compute-sanitizer (racecheck, memcheck) over the parity tests that drive the mix kernel's lean batches (uniform batches of
# whole-tile cells: cfg 2 / cfg 3 shapes, sharded sessions, page-locked output), its general loop (golden scenarios, fuzz,
# ragged blocks) and fx_chain_kernel's role table.   gpurun -- bash tools/gpu_sanitizer_mix.sh
set -x
O=gpurun_out/sanitizer
mkdir -p $O
SEL="golden_exact or golden_fuzz or golden_tree or every_tile_shape or cfg2_256 or cfg3_128 or cfg1_mono or block_sizes or unaligned or page_locked or sharded_peer or sharded_rank or golden_sharded or effects_every_kernel_shape or effects_bench_shape or polyphase"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/compute_sanitizer_racecheck.log
tail -4 $O/compute_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/compute_sanitizer_memcheck.log
tail -3 $O/compute_sanitizer_memcheck.log
