python scratch/diag_b.py 2>&1 | tail -16
bash scratch/run_t1.sh
bash scratch/ncu_p.sh > /dev/null 2>&1
