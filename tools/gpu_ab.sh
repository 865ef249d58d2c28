# A/B on the GPU box: the GPU test suite, then bench stage times under tuning environments / library variants
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/variants.sh base minb2 coop2
FGL_LIB=fauxgl_b200/libfauxgl_b200.coop2.so FGL_TILE_CLOCK=1 python tools/tile_cycles.py 2>&1 | sed -n 2,5p
