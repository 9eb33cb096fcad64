"""Material constants of the four species the reference knows (MC/tools/material.py:151-302) and the
derived densities/energies (material.py:15-88).  Same attribute names as the reference's ``material``
class (species, params, atnd, plasmon_e, fermi_e) so either object can be handed to the workers."""
import math

from scipy.constants import Avogadro, hbar, m_e, e as _q, epsilon_0

_EV = 1.602176634e-19

_SPECIES = {
    'Al': dict(Z=13, density=2.70, atwt=26.98154, n_val=3, E_val=72.55, name_val='3s2-3p1',
               name_s=['1s', '2s', '2p'], Es={'1s': 1559., '2s': 118., '2p': 73.5},
               ns={'1s': 2, '2s': 2, '2p': 6}, bethe_k=0.815, penn_b=-2.16),
    'Si': dict(Z=14, density=2.33, atwt=28.085, n_val=4, E_val=99.42, name_val='3s2-3p2',
               name_s=['1s', '2s', '2p'], Es={'1s': 1189, '2s': 149, '2p': 100},
               ns={'1s': 2, '2s': 2, '2p': 6}, bethe_k=0.822, penn_b=-2.19, xip_g=1360),
    'Cu': dict(Z=29, density=8.96, atwt=63.546, n_val=11, E_val=75.1, name_val='3d10-4s2',
               name_s=['1s', '2s2p', '3s', '3p'], Es={'1s': 8980, '2s2p': 977, '3s': 120, '3p': 74},
               ns={'1s': 2, '2s2p': 8, '3s': 2, '3p': 6}, bethe_k=0.83, penn_b=-3.21),
    'Au': dict(Z=79, density=19.30, atwt=196.967, n_val=11, E_val=75.1, name_val='5d10-6s1',
               name_s=['2s2p', '3s3p3d', '4s4p', '4d', '5s', '4f', '5p2', '5p4'],
               Es={'2s2p': 12980, '3s3p3d': 2584, '4s4p': 624, '4d': 341, '5s': 178, '4f': 85, '5p2': 72, '5p4': 54},
               ns={'2s2p': 8, '3s3p3d': 18, '4s4p': 8, '4d': 10, '5s': 2, '4f': 14, '5p2': 2, '5p4': 4},
               bethe_k=0.851, penn_b=-2.16),
}


def scatter_params(species):
    """material.py:151-302."""
    if species not in _SPECIES:
        raise KeyError('unknown species %r (the reference defines Al, Si, Cu, Au)' % (species,))
    d = dict(_SPECIES[species])
    d['species'] = species
    d['name_s'] = list(d['name_s'])
    d['Es'] = dict(d['Es'])
    d['ns'] = dict(d['ns'])
    return d


def at_num_dens(dens, atom_mass):
    """material.py:15-30: atoms per m^3."""
    return dens * Avogadro / atom_mass / (1e-2 ** 3)


def fermi_energy(atNumDens, nvalence):
    """material.py:34-56, eV."""
    n = nvalence * atNumDens
    return hbar ** 2 * (3. * (math.pi ** 2) * n) ** (2. / 3.) / (2. * m_e) / _EV


def plasmon_energy(atNumDens, nvalence):
    """material.py:61-88, eV."""
    n = nvalence * atNumDens
    return hbar * _q * (n / (epsilon_0 * m_e)) ** 0.5 / _EV


class material:
    """material.py:117-146."""
    __slots__ = ['species', 'params', 'atnd', 'plasmon_e', 'fermi_e']

    def __init__(self, species):
        self.species = species
        self.params = scatter_params(species)
        self.atnd = at_num_dens(self.params['density'], self.params['atwt'])
        self.plasmon_e = plasmon_energy(self.atnd, self.params['n_val'])
        self.fermi_e = fermi_energy(self.atnd, self.params['n_val'])

    @property
    def channel_names(self):
        """sigma-dict insertion order (scattering.py:485-508) = the `type` labels of scatter_output."""
        return ['Ruth', 'Moller'] + ['Gryz' + s for s in self.params['name_s']] + ['Quinn']


def as_c_material(mat):
    """Fill the C struct from any object with the reference's material attributes."""
    from ._lib import Material, MAX_SHELLS
    P = mat.params
    m = Material()
    m.Z = int(P['Z'])
    names = list(P['name_s'])
    if not 1 <= len(names) <= MAX_SHELLS:
        raise ValueError('between 1 and %d core shells supported' % MAX_SHELLS)
    m.n_shells = len(names)
    m.n_val = float(P['n_val'])
    m.E_val = float(P['E_val'])
    m.atnd = float(mat.atnd)
    m.fermi_e = float(mat.fermi_e)
    m.plasmon_e = float(mat.plasmon_e)
    m.bethe_k = float(P['bethe_k'])
    for i, s in enumerate(names):
        m.shell_E[i] = float(P['Es'][s])
        m.shell_n[i] = float(P['ns'][s])
    return m
