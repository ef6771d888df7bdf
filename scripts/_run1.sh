timeout 900 python -m pytest tests/test_sharded.py -x -q -m gpu 2>&1 | tail -5
bash scripts/_run8.sh ${1:-2}
