#!/bin/bash
# phase line of the batched contacts call, N runs (GPU box): tools/diag/contacts_ab.sh [runs] [npose]
for i in $(seq ${1:-3}); do B2C_TIMING=1 python tools/contacts_bench.py ${2:-16384} 2>&1 | grep contacts_batch | tail -1 | cut -c40-330; done
