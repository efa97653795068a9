run() { timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['parity_sample_ok'])"; }
for l2 in 16 24 32 48 64; do echo "N=1000 l2=$l2 MB"; IGMK_L2_BUDGET=$((l2*1048576)) run; done
for l2 in 24 48; do echo "N=10000 l2=$l2 MB"; IGMK_L2_BUDGET=$((l2*1048576)) run --nstruct 10000 --max-pairs 400000; done
for sl in 1048576 2097152; do echo "e2e slice=$sl"; IGMK_HOST_SLICE=$sl timeout 300 python bench.py --no-cpu-baseline --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'])"; done
