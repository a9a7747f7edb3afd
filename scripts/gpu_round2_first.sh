# First GPU call of the next round (1 GPU, ~6 min): confirm the committed engine, then time the queued variants.
#   here (CPU):  bash scripts/build_variants.sh
#   then:        gpurun --timeout 600 -- 'bash scripts/gpu_round2_first.sh'
# A variant that wins must pass   MOPS_B200_LIB=$PWD/build_variants/<name>.so python -m pytest tests -m gpu -q   before its
# flag becomes the default.  Multi-GPU follow-up: gpurun --gpus 2 -- 'bash scripts/gpu_scale2_allgather.sh'.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash scripts/gpu_ab.sh
for f in build_variants/base.so build_variants/sq_filter.so; do
  [ -f $f ] && MOPS_B200_LIB=$PWD/$f timeout 150 python scripts/bench_secondary.py 2>/dev/null | grep "C3 streamline" | cut -c1-260
done
