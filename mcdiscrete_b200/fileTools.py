"""input.file parser and the output column contract (MC/tools/fileTools.py:8-46, 67-99, 193-242)."""

STR_KEYS = ('mode', 'material', 'elastic', 'Bethe_model')
INT_KEYS = ('num_el', 'maxScatt', 'seed')
BOOL_KEYS = ('gen_tables', 'diffMFP')

EL_COLUMNS = ('outcome', 'final_E', 'final_dir', 'last_pos')
SCAT_COLUMNS = ('type', 'position', 'pathl', 'E', 'E_loss', 'pol_angle', 'az_angle', 'energy', 'direction')
_NESTED = {'final_dir': ('dx', 'dy', 'dz'), 'position': ('x', 'y', 'z'), 'last_pos': ('last_x', 'last_y', 'last_z')}


def readInput(fileName='input.file'):
    """fileTools.py:8-46: ``key : value  # comment`` lines."""
    data = {}
    with open(fileName) as fh:
        for line in fh.read().split('\n'):
            if not line.strip() or line.startswith('#'):
                continue
            line = line.partition('#')[0]
            param = [p.strip(' ') for p in line.split(':')]
            key = param[0]
            if key in STR_KEYS:
                data[key] = param[1]
            elif key in INT_KEYS:
                data[key] = int(param[1])
            elif 'output' in key:
                data[key] = param[1].replace(' ', '').split(',')
            elif key in BOOL_KEYS:
                data[key] = 'es' in param[1]
            else:
                data[key] = float(param[1])
    return data


class thingsToSave:
    """fileTools.py:193-242: only requested names are recorded; the three vector-valued names are nested
    dicts of scalar columns."""

    def __init__(self, args):
        self.dict = {}
        for arg in args:
            if arg in _NESTED:
                self.dict[arg] = {k: [] for k in _NESTED[arg]}
            else:
                self.dict[arg] = []

    def addToList(self, par, value):
        if par in self.dict:
            if par in _NESTED:
                for k, v in zip(_NESTED[par], value):
                    self.dict[par][k].append(v)
            else:
                self.dict[par].append(value)


def zipDict(dictA, dictB):
    """fileTools.py:67-99: flatten final_dir / position / last_pos and extend A's lists with B's."""
    for nested in ('final_dir', 'position', 'last_pos'):
        if nested in dictB:
            for k in dictB[nested]:
                dictB[k] = dictB[nested][k]
            del dictB[nested]
    for k in dictB:
        if k in dictA:
            dictA[k] += dictB[k]
        else:
            dictA[k] = dictB[k]
    return dictA


def flat_columns(requested, which):
    """Column order after zipDict (fileTools.py:72-91): scalars in requested order, then the flattened
    vector columns in the order final_dir, position, last_pos."""
    known = EL_COLUMNS if which == 'els' else SCAT_COLUMNS
    cols = [c for c in requested if c in known and c not in _NESTED]
    for nested in ('final_dir', 'position', 'last_pos'):
        if nested in requested and nested in known:
            cols += list(_NESTED[nested])
    if which == 'scats' and 'direction' in requested:
        pass            # 'direction' stays one object column in the reference (electron.py:99)
    return cols
