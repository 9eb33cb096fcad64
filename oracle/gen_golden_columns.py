"""TEST INFRASTRUCTURE ONLY.  The reference's own thingsToSave / zipDict (MC/tools/fileTools.py:67-99, 193-242) and
the DataFrame column order its listener stores (doScatter.py:99-119), driven through a scripted sequence of calls
-> tests/golden/columns.json.   python oracle/gen_golden_columns.py"""
import copy
import itertools
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

ROOT = os.path.dirname(HERE)
EL = ['outcome', 'final_E', 'final_dir', 'last_pos']
SC = ['type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy']


def script(T, names):
    """two 'electrons' worth of addToList calls, including names that were not requested"""
    out = []
    for el in range(2):
        t = T(names)
        for ev in range(2 + el):
            t.addToList('type', 'Ruth' if ev % 2 == 0 else 'Gryz2p')
            t.addToList('position', [1.0 * el, 2.0 * ev, 3.0])
            t.addToList('pathl', 10.0 + ev)
            t.addToList('E', 20000.0 - ev)
            t.addToList('E_loss', 16.5 * ev)
            t.addToList('pol_angle', 0.99)
            t.addToList('az_angle', 0.25)
            t.addToList('energy', 19990.0 - ev)
            t.addToList('outcome', 'bks')
            t.addToList('final_E', 12345.0)
            t.addToList('final_dir', [0.1, 0.2, -0.9])
            t.addToList('last_pos', [5.0, 6.0, 7.0])
        out.append(t.dict)
    return out


def main():
    import ref_shim
    ref_shim.load_reference()
    from tools import fileTools as ft
    import pandas as pd
    cases = {}
    subsets = [EL, ['outcome'], ['final_dir', 'outcome'], ['last_pos', 'final_E', 'final_dir'], SC, ['energy', 'type'],
               ['position', 'type', 'energy'], ['E_loss', 'position'], []]
    for names in subsets:
        dicts = script(ft.thingsToSave, names)
        merged = {}
        for d in copy.deepcopy(dicts):
            merged = ft.zipDict(merged, d)
        cols = list(pd.DataFrame.from_dict(merged).columns) if merged else []
        cases[','.join(names)] = {'requested': names, 'per_electron': dicts, 'zipped': merged, 'frame_columns': cols}
    json.dump(cases, open(os.path.join(ROOT, 'tests', 'golden', 'columns.json'), 'w'), indent=1)
    print(len(cases), 'cases;', cases[','.join(subsets[6])]['frame_columns'])


if __name__ == '__main__':
    main()
