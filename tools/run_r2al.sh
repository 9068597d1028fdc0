set -x
for lg in 17 18 19 20 21; do FORGE_EC_CHUNK_LG=$lg timeout 120 python tools/sweep_e2e.py 2>&1 | tail -1; done
