#!/usr/bin/env bash
{
timeout 300 python scripts/perf.py --scene cornell
timeout 300 python scripts/perf.py --scene cornell --opt trace_ctas_per_sm=5
timeout 300 python scripts/perf.py --scene cornell --opt trace_ctas_per_sm=7
timeout 300 python scripts/perf.py --scene vol --size 512
timeout 300 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --tag c1
} 2>&1 | grep -E "PERF|rror"
