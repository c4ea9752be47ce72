set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
j() { grep "^{" | tail -1; }
# strong scaling, small N, concurrently on disjoint GPUs
( CUDA_VISIBLE_DEVICES=0,1,2,3 $TR --nproc-per-node 4 --master-port 29511 bench.py --gpus 4 --config c3 --quick --steps 10 --warmup 3 2>gpurun_out/r02g_c3_n4.err | j > gpurun_out/r02g_c3_n4.json ) &
( CUDA_VISIBLE_DEVICES=4,5 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --config c3 --quick --steps 10 --warmup 3 2>gpurun_out/r02g_c3_n2.err | j > gpurun_out/r02g_c3_n2.json ) &
( CUDA_VISIBLE_DEVICES=6 python bench.py --config c3 --quick --steps 10 --warmup 3 2>gpurun_out/r02g_c3_n1.err | j > gpurun_out/r02g_c3_n1.json ) &
( CUDA_VISIBLE_DEVICES=7 python bench.py --config c4 --quick --steps 10 --warmup 3 2>gpurun_out/r02g_c4_n1.err | j > gpurun_out/r02g_c4_n1.json ) &
wait
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --config c3 --quick --steps 10 --warmup 3 2>gpurun_out/r02g_c3_n8.err | j > gpurun_out/r02g_c3_n8.json
$TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --config c4 --quick --steps 10 --warmup 3 2>gpurun_out/r02g_c4_n8.err | j > gpurun_out/r02g_c4_n8.json
$TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 30 --warmup 5 --no-train --no-cpu-baseline --parity-images 32 2>gpurun_out/r02g_c2_n8.err | j > gpurun_out/r02g_c2_n8.json
$TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --steps 30 --warmup 5 --quick --no-balance 2>gpurun_out/r02g_c2_n8_eq.err | j > gpurun_out/r02g_c2_n8_eq.json
python bench.py --steps 30 --warmup 5 --quick 2>gpurun_out/r02g_c2_n1.err | j > gpurun_out/r02g_c2_n1.json
ls -la gpurun_out/r02g_*
tail -2 gpurun_out/r02g_*.err | cut -c1-200
