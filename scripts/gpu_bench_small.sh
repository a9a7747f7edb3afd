set -x
python __graft_entry__.py smoke 2>&1 | tail -5
python bench.py --level 6 --layers 20 --particles 400000 --steps 3 --warmup 3 --cpu-level 5 2>&1 | tail -5
