export SLR_B200_LIB=$PWD/structure-light-reconstructor_b200/libslr_b200_abl.so
b() { timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"; }
for a in 0 1 2 3 4 8 12 16 31; do echo -n "ablate $a: "; SLR_ABLATE=$a b; done
