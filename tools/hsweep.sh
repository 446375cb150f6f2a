for H in 256 1024 2048 4096; do python tools/prof_fused.py $H 4096 256 3 1,2 2>&1 | grep -E "fused|identical"; done
