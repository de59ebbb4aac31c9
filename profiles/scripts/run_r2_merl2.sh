sed -n '/^cat > \/tmp\/merl_t.py/,/^PY$/p' profiles/scripts/run_r2_merl.sh > /tmp/mk.sh; bash /tmp/mk.sh
python /tmp/merl_t.py
DJB200_MERL_NOCARVE=1 python /tmp/merl_t.py
DJB200_MERL_PROBE=2 python /tmp/merl_t.py
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -k merl 2>&1 | tail -2
