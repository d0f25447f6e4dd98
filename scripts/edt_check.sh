python -m pytest tests/test_edt_build.py -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_edt" --csv --log-file gpurun_out/s15_qp.csv python scripts/edt_probe.py > /dev/null 2>&1
python scripts/ncu_print.py | tail -9
python scripts/edt_probe.py time
