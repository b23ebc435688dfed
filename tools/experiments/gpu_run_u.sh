#!/bin/bash
gcc -O3 -march=x86-64-v3 -pthread tools/experiments/membw.c -o /tmp/membw && for t in 1 16 64 128; do /tmp/membw $t 1000000000; done
nproc; lscpu | grep -E "Model name|Socket|NUMA node|Thread|MHz" | head -8; numactl -H 2>/dev/null | head -6; cat /sys/kernel/mm/transparent_hugepage/enabled
