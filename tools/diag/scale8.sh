# one 8-GPU box: N=1 reference, then N=8 with the peer exchange, the NCCL exchange, and config 5 sharded over 8
set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/scale_n8_peer.json 2> gpurun_out/scale_n8_peer.err
timeout 300 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-sub --exchange nccl > gpurun_out/scale_n8_nccl.json 2> gpurun_out/scale_n8_nccl.err
timeout 300 $TR --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-sub --workload config5 > gpurun_out/scale_n8_config5.json 2> gpurun_out/scale_n8_config5.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub --workload config5 > gpurun_out/scale_n1_config5.json 2> gpurun_out/scale_n1_config5.err
tail -c 300 gpurun_out/scale_n8_peer.err
