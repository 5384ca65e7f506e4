timeout 100 python tools/microbench.py 512 512x512x64 --ring-only 2>&1 | grep -o "n=.*us/it = [0-9]* GB/s"
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
