timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -3
GCM_B200_OVERLAP=0 timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -1
bash tools/scale_run.sh 2
GCM_B200_OVERLAP=0 bash tools/scale_run.sh 2
