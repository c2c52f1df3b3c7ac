#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench exit $?"
tail -5 gpurun_out/bench14.err; cat gpurun_out/bench14.json
