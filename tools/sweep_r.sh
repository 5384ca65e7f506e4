timeout 100 python tools/microbench.py 512 512x512x64 256x256x512 --ring-only 2>&1 | grep -o "n=.*us/it = [0-9]* GB/s"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cg_ring -c 1 -o gpurun_out/cg_r1b -f python tools/microbench.py 512 --ring-only > gpurun_out/ncu_cg_r1b.log 2>&1; tail -2 gpurun_out/ncu_cg_r1b.log
