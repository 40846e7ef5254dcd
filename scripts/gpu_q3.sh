#!/usr/bin/env bash
{
timeout 300 python scripts/perf.py --scene cornell
timeout 300 python scripts/perf.py --scene cornell --opt small_coop=0
timeout 300 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --tag c1
timeout 300 python scripts/perf.py --scene cornell4 --size 256 --spp 64 --tag c1 --opt small_coop=0
} 2>&1 | grep -E "PERF|rror"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
