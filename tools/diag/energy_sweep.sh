# A/B of energy-kernel build / run-time knobs: PRESET step, 20 timed steps each
for n in base mb3 mb5 lt24 lt16; do
  L=$PWD/graspit_b200/libb2c_$n.so; [ $n = base ] && L=$PWD/graspit_b200/libb2c.so
  B2C_LIB=$L python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/sw_$n.json 2> gpurun_out/sw_$n.err
done
for s in 0 32 512; do B2C_STAGE_NODES=$s python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/sw_stage$s.json 2> gpurun_out/sw_stage$s.err; done
for s in 0 32 128; do B2C_SEED_RES=$s python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/sw_seed$s.json 2> gpurun_out/sw_seed$s.err; done
