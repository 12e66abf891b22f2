#!/bin/bash
for u in 2 4 8; do
  echo "unroll $u: $(SCG_RANDOM_UNROLL=$u python tools/profile_config.py 5 200000000 3 2>&1 | tail -1 | grep -o '[0-9.]* ms per pass.*')"
done
for p in 8 16 32; do
  echo "parts $p: $(SCG_RANDOM_PARTS=$p python tools/profile_config.py 5 200000000 3 2>&1 | tail -1 | grep -o '[0-9.]* ms per pass.*')"
done
