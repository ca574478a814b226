#!/bin/bash
timeout 300 python tools/chain_bench.py 16
PC_CHAIN_FLAGS=16 timeout 300 python tools/chain_bench.py 16
PC_CHAIN_FLAGS=20 timeout 300 python tools/chain_bench.py 16
