"""TEST INFRASTRUCTURE ONLY.  What the reference's own readInput (MC/tools/fileTools.py:8-46) returns for the input
files under tests/data/ -> tests/golden/readInput.json.   python oracle/gen_golden_input.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim                      # noqa: E402

ROOT = os.path.dirname(HERE)


def main():
    ref_shim.load_reference()
    from tools import fileTools as ft
    out = {}
    for name in ('input.file', 'input_variants.file'):
        out[name] = ft.readInput(os.path.join(ROOT, 'tests', 'data', name))
    json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'readInput.json'), 'w'), indent=1)
    print(json.dumps(out['input_variants.file']))


if __name__ == '__main__':
    main()
