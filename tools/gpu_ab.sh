# A/B on the GPU box: the GPU test suite, then bench stage times of library variants
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/variants.sh base
