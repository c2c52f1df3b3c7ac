#!/bin/bash
timeout 300 python tools/trace_step.py 2>&1 | tail -40
