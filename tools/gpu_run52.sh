#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --maxfail=8 -p no:cacheprovider 2>&1 | tail -5
