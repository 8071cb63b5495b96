python -m pytest tests -m gpu -x -q -k "mods_pairs or mods_pair_with_mser" 2>&1 | tail -3
python tools/e2e_probe.py 12 2 2>&1 | tail -20
