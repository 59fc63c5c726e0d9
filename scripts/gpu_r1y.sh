timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python scripts/cfg4_probe.py 2>&1 | grep -i "^fisher\|^z \|^loo \|z mean"
bash scripts/gpu_profile_all.sh r1g "loo_em_step5 loo_like2 fisher2 em_pop_multi2" 2>&1 | grep "rc="
