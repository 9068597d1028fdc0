echo "== default"; python tools/sweep_small.py 2>&1 | tail -6
echo "== FORGE_EC_FIXED_IMPL=vm"; FORGE_EC_FIXED_IMPL=vm python tools/sweep_small.py 2>&1 | tail -6
echo "== FORGE_EC_FIXED_IMPL=regs"; FORGE_EC_FIXED_IMPL=regs python tools/sweep_small.py 2>&1 | tail -6
