"""TEST INFRASTRUCTURE ONLY -- CPU restatement (float64, numpy / pure Python) of the
per-electron trajectory loop of elena-pascal/MC-discrete.

Nothing under ``mcdiscrete_b200/`` imports this file.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
use it, and only as the checker / the CPU baseline -- never as the thing shipped.

Parity status: PINNED.  Every function below is checked against outputs of the
reference's own unmodified code (imported under ``oracle/ref_shim.py`` in the build
container, vectors committed in ``tests/golden/`` by ``oracle/gen_golden.py``); see
``tests/test_oracle_golden.py``.  The reference ships no golden vectors of its own
(SURVEY.md section 8(c)).

All ``file:line`` citations are relative to the reference checkout (``MC/tools/...``).
Units follow the reference: eV, Angstrom, cm^2, m^-3.
"""
import bisect
import math
import os

import numpy as np

# --------------------------------------------------------------------------------------
# constants (parameters.py:7-8; scipy.constants CODATA values used by material.py:9)
# --------------------------------------------------------------------------------------
PI_EFOUR = 6.51408491409531e-14          # parameters.py:7   cm^2 eV^2
try:
    from scipy.constants import physical_constants, hbar, m_e, e as q_e, epsilon_0, Avogadro
    BOHR_R = physical_constants['Bohr radius'][0]   # parameters.py:8
except Exception:                                    # pragma: no cover
    BOHR_R, hbar, m_e, q_e, epsilon_0, Avogadro = (5.29177210544e-11, 1.054571817e-34,
                                                   9.1093837139e-31, 1.602176634e-19,
                                                   8.8541878188e-12, 6.02214076e23)
EV_J = 1.602176634e-19

OUTCOMES = ('abs', 'bks', 'trsm', 'trdf', 'long')


# --------------------------------------------------------------------------------------
# material (material.py:15-88, 117-146, 151-302)
# --------------------------------------------------------------------------------------
_SPECIES = {
    # species: Z, density, atwt, n_val, E_val, name_val, [(shell, Es, ns)...], bethe_k
    'Al': (13, 2.70, 26.98154, 3, 72.55, '3s2-3p1',
           [('1s', 1559., 2), ('2s', 118., 2), ('2p', 73.5, 6)], 0.815),
    'Si': (14, 2.33, 28.085, 4, 99.42, '3s2-3p2',
           [('1s', 1189, 2), ('2s', 149, 2), ('2p', 100, 6)], 0.822),
    'Cu': (29, 8.96, 63.546, 11, 75.1, '3d10-4s2',
           [('1s', 8980, 2), ('2s2p', 977, 8), ('3s', 120, 2), ('3p', 74, 6)], 0.83),
    'Au': (79, 19.30, 196.967, 11, 75.1, '5d10-6s1',
           [('2s2p', 12980, 8), ('3s3p3d', 2584, 18), ('4s4p', 624, 8), ('4d', 341, 10),
            ('5s', 178, 2), ('4f', 85, 14), ('5p2', 72, 2), ('5p4', 54, 4)], 0.851),
}


class Material:
    """material.py:117-146 with scatter_params :151-302."""

    def __init__(self, species):
        Z, dens, atwt, n_val, E_val, name_val, shells, k = _SPECIES[species]
        self.species = species
        self.Z, self.density, self.atwt = Z, dens, atwt
        self.n_val, self.E_val, self.name_val = n_val, E_val, name_val
        self.name_s = [s[0] for s in shells]
        self.Es = [float(s[1]) for s in shells]
        self.ns = [float(s[2]) for s in shells]
        self.bethe_k = k
        self.atnd = dens * Avogadro / atwt / (1e-2 ** 3)                  # material.py:30 (m**3/cm**3)
        n = n_val * self.atnd
        self.fermi_e = hbar ** 2 * (3. * (math.pi ** 2) * n) ** (2. / 3.) / (2. * m_e) / EV_J   # :54-56
        self.plasmon_e = hbar * q_e * (n / (epsilon_0 * m_e)) ** 0.5 / EV_J                      # :86-88

    @property
    def channel_names(self):
        """sigma-dict insertion order (scattering.py:485-508)."""
        return ['Ruth', 'Moller'] + ['Gryz' + s for s in self.name_s] + ['Quinn']


# --------------------------------------------------------------------------------------
# cross sections / stopping powers
# --------------------------------------------------------------------------------------
def alpha(E, Z):
    """crossSections.py:14-27."""
    return 3.4 * (Z ** 0.666667) / E


def ruther_sigma(E, Z):
    """crossSections.py:37-58."""
    a = alpha(E, Z)
    Ek = E * 1e-3
    return 5.21e-21 * (Z * Z / (Ek * Ek)) * ((Ek + 511.) / (Ek + 1024.)) ** 2 * (4. * math.pi) / (a * (1. + a))


def ruther_sigma_wDefl(E, Z):
    """crossSections.py:63-84: Z**2 -> Z*(Z+1)."""
    a = alpha(E, Z)
    Ek = E * 1e-3
    return 5.21e-21 * (Z * (Z + 1) / (Ek * Ek)) * ((Ek + 511.) / (Ek + 1024.)) ** 2 * (4. * math.pi) / (a * (1. + a))


def ruther_N_sigma(E, Z):
    """crossSections.py:89-112 (Nigam screening, beta* = 0.48 beta_N)."""
    b = 0.48 * (5.43 * Z ** (2 / 3) / E)
    return PI_EFOUR * Z ** 2 / (4 * b * (1 + b) * E ** 2)


def ruther_N_sigma_wDefl(E, Z):
    """crossSections.py:115-143."""
    b = 0.48 * (5.43 * Z ** (2 / 3) / E)
    return PI_EFOUR * (Z + 1) * Z / (4 * b * (1 + b) * E ** 2)


ELASTIC_MODELS = ('Ruth_vanilla', 'Ruth_vanilla_wDefl', 'Ruth_nigram', 'Ruth_nigram_wDefl')


def elastic_sigma(E, Z, elastic='Ruth_vanilla', mode='DS'):
    """The elastic cross-section the reference evaluates for an `elastic` label.
    DS (scattering.py:483-491): exact label match / 'vanilla_wDefl' / 'nigram_wDefl'.  'Ruth_vanilla_wDefl'
    raises NameError there (ruther_sigma_wDefl is not imported, scattering.py:14); the intended function is used.
    Continuous (scattering.py:781-789): substring tests in the order 'vanilla', 'vanilla_wDefl', 'nigram',
    'nigram_wDefl', so the two _wDefl branches are unreachable and the labels fall back to the plain forms."""
    if elastic not in ELASTIC_MODELS:
        raise ValueError('elastic model %r is not on the path' % (elastic,))
    if mode != 'DS':
        return ruther_sigma(E, Z) if 'vanilla' in elastic else ruther_N_sigma(E, Z)
    return {'Ruth_vanilla': ruther_sigma, 'Ruth_vanilla_wDefl': ruther_sigma_wDefl,
            'Ruth_nigram': ruther_N_sigma, 'Ruth_nigram_wDefl': ruther_N_sigma_wDefl}[elastic](E, Z)


def moller_sigma(E, Wc, nfree):
    """crossSections.py:183-208."""
    if E <= Wc:
        return 0.
    eps = Wc * 1. / E
    return nfree * PI_EFOUR * (1. / (E * E)) * (1. / eps - 1. / (1. - eps) + math.log(eps / (1. - eps)))


def gryz_sigma(E, Esi, nsi):
    """crossSections.py:213-242."""
    if E <= Esi:
        return 0.
    U = E * 1. / Esi
    return nsi * PI_EFOUR * ((U - 1.) / (U + 1.)) ** 1.5 * \
        (1. + (2. * (1. - (1. / (2. * U))) / 3. * math.log(2.7 + ((U - 1.) ** 0.5)))) / (E * Esi)


def quinn_sigma(E, Epl, Ef, n):
    """crossSections.py:247-284 (E <= Epl raises in the reference; unreachable for Emin >> Epl)."""
    E_Ef = E * 1. / Ef
    Epl_Ef = Epl * 1. / Ef
    tot = Epl * math.log(((1. + Epl_Ef) ** 0.5 - 1.) / (E_Ef ** 0.5 - (E_Ef - Epl_Ef) ** 0.5)) / (2. * BOHR_R * E)
    return (tot / n) * 1e4


def mfp_from_sigma(sigma, n):
    """scattering.py:20-36:  1/(n sigma) m^3/cm^2/Angstrom = 1e14/(n sigma)."""
    return 1. / (n * sigma) * 1e14


class MottTable:
    """extFunctions.py:21-109: numerical elastic cross-sections in the ioffe.ru/ES/Elastic layout.
    ``Es`` ascending (the file is descending and is flipped, :93-97), ``sigmas[i]`` in cm^2, ``probTable[i][j]`` =
    cumulative SUM of dCS over the angle grid divided by the row sum (no solid-angle weight, :99-109)."""

    def __init__(self, Z, name, path='../tables/mott/'):
        self.Z, self.name = Z, name
        self.ioffeFile = os.path.join(path + 'ioffe/' + name + '.ioffe')                  # :47
        with open(self.ioffeFile, 'r') as f:
            lines = f.readlines()
        if int(lines[2].split()[-1]) != Z:                                                  # :58-60
            raise ValueError('Trying to read file %s but expecting Z:%s' % (self.ioffeFile, Z))
        self.sizeTheta = int(lines[5].split()[0])                                           # :62
        self.sizeE = int(lines[6].split()[0])                                               # :63
        self.Es = np.empty(self.sizeE)
        self.sigmas = np.empty(self.sizeE)
        self.dCS = np.empty([self.sizeE, self.sizeTheta])
        self.thetas = None
        self.probTable = None

    def read_ioffe(self):                                                                   # :75-97
        with open(self.ioffeFile, 'r') as f:
            lines = f.readlines()
        self.thetas = np.array([float(x) for x in lines[12].strip('theta [grad]=').split()])
        for i, line in enumerate(lines[14:-1]):
            items = line.replace('|', '').split()
            self.Es[i] = float(items[0])
            self.sigmas[i] = float(items[-1])
            self.dCS[i] = np.array([float(x) for x in items[1:-1]])
        self.Es = np.flip(self.Es)
        self.sigmas = np.flip(self.sigmas)
        self.dCS = np.flip(self.dCS, axis=0)
        return self

    def to_prob_table(self):                                                                # :99-109
        cum = np.cumsum(self.dCS, axis=1)
        self.probTable = cum / np.broadcast_to(np.sum(self.dCS, axis=1), cum.T.shape).T
        return self


def mott_sigma(tab, E):
    """scattering.py:493: sigmas[bisect_left(Es, E)] (cm^2)."""
    return float(tab.sigmas[bisect.bisect_left(tab.Es, E)])


def pick_mott_table(tab, E, u):
    """scattering.py:96-122: row bisect_left(Es, E), column bisect_left(CDF row, u), cos^2(radians(theta/2)).
    Returns (c2, iE, jT)."""
    iE = bisect.bisect_left(tab.Es, E)
    jT = bisect.bisect_left(tab.probTable[iE], u)
    theta = tab.thetas[jT]
    return float(np.cos(np.radians(theta * 0.5)) ** 2), iE, jT


def sigmas_ds(E, mat, Wc, elastic='Ruth_vanilla', mott=None):
    """scattering.py:485-508: list in insertion order Ruth (or Mott, :493), Moller, Gryz<shell>..., Quinn."""
    s = [mott_sigma(mott, E) if elastic == 'Mott' else elastic_sigma(E, mat.Z, elastic, 'DS'),
         moller_sigma(E, Wc, mat.n_val)]
    s += [gryz_sigma(E, mat.Es[i], mat.ns[i]) for i in range(len(mat.Es))]
    s.append(quinn_sigma(E, mat.plasmon_e, mat.fermi_e, mat.atnd))
    return s


def mean_ionisation_J(Z):
    """stoppingPowers.py:134-137."""
    return (9.76 * Z + 58.5 / (Z ** 0.19)) if Z >= 13 else 11.5 * Z


def bethe_cl_sp(Z, E, n):
    """stoppingPowers.py:112-140 (eV/Angstrom)."""
    J = mean_ionisation_J(Z)
    return 2. * PI_EFOUR * n * (Z / E) * math.log(1.166 * E / J) * 1e-14


def bethe_mod_sp_k(Z, E, n, k):
    """stoppingPowers.py:145-173 (Joy-Luo)."""
    J = mean_ionisation_J(Z)
    return 2. * PI_EFOUR * n * (Z / E) * math.log(1.166 * (E + k * J) / J) * 1e-14


def bethe_mod_sp(E, n, Zi, Ei, Zval, Eval):
    """stoppingPowers.py:177-215 (explicit shells)."""
    pref = (2. * PI_EFOUR * n / E) * 1e-14
    s = Zval * math.log(E * 1. / Eval)
    for z, eb in zip(Zi, Ei):
        s = s + z * math.log(E * 1. / eb)
    return pref * s


def stopping_power(model, E, mat):
    """Bethe model dispatch as intended by multiScatter.py:205-215 / scattering.py:897-913
    (the shipped non-unit JL / explicit classes call the classical form: SURVEY A23/A24)."""
    if model == 'classical':
        return bethe_cl_sp(mat.Z, E, mat.atnd)
    if model == 'JL':
        return bethe_mod_sp_k(mat.Z, E, mat.atnd, mat.bethe_k)
    if model == 'explicit':
        return bethe_mod_sp(E, mat.atnd, mat.ns, mat.Es, mat.n_val, mat.E_val)
    raise ValueError(model)


# --------------------------------------------------------------------------------------
# channel pick / angles
# --------------------------------------------------------------------------------------
def pick_from_sigmas(sig, u):
    """scattering.py:39-59: stable descending sort, cumsum/sum, bisect_left(u).
    Returns the index into ``sig`` (insertion order)."""
    order = sorted(range(len(sig)), key=lambda i: sig[i], reverse=True)   # stable, like sorted(..., reverse=True)
    vals = [sig[i] for i in order]
    probs = np.cumsum(vals) / sum(vals)
    k = bisect.bisect_left(probs, u)
    return order[k]


def rutherford_halfpol(E, Z, u):
    """scattering.py:125-152: cos^2(theta/2)."""
    a = alpha(E, Z)
    return 1. - (a * u / (1. + a - u))


def binary_coll(E, W, Ef):
    """scattering.py:156-179 with probTables.py:17-25."""
    if W <= (E - Ef) * 0.5:
        return 0.5 * (math.sqrt(1 - (W / E)) + 1)
    return float('nan')


def newdir(s_hT, c_hT, s_hP, c_hP, y_local, d):
    """rotation.py:41-53 with numpy-quaternion's Hamilton product and rotate_vectors."""
    pw, pv = c_hT, np.asarray(y_local, float) * s_hT
    aw, av = c_hP, np.asarray(d, float) * s_hP
    w = aw * pw - av @ pv                                   # q_total = q_az * q_polar
    v = aw * pv + pw * av + np.cross(av, pv)
    x, y, z = v
    n = w * w + x * x + y * y + z * z
    m = np.array([
        [1 - 2 * (y * y + z * z) / n, 2 * (x * y - z * w) / n, 2 * (x * z + y * w) / n],
        [2 * (x * y + z * w) / n, 1 - 2 * (x * x + z * z) / n, 2 * (y * z - x * w) / n],
        [2 * (x * z - y * w) / n, 2 * (y * z + x * w) / n, 1 - 2 * (x * x + y * y) / n]])
    return m @ np.asarray(d, float), m @ np.asarray(y_local, float)


def update_direction(d, y_local, c2, halfPhi):
    """electron.py:87-99."""
    s_hT = (1. - c2) ** 0.5
    c_hT = c2 ** 0.5
    s_hP = math.sin(halfPhi)
    c_hP = (1. - s_hP ** 2) ** 0.5
    nd, ny = newdir(s_hT, c_hT, s_hP, c_hP, y_local, d)
    return nd / np.linalg.norm(nd), ny / np.linalg.norm(ny)


# --------------------------------------------------------------------------------------
# tables (probTables.py, extFunctions.py)
# --------------------------------------------------------------------------------------
def moller_dCS(E, W, nfree):
    """extFunctions.py:159-184 (dtype follows the inputs, as in the reference)."""
    eps = W / E
    with np.errstate(all='ignore'):
        return np.where(eps < 1., nfree * PI_EFOUR * (1. / (eps ** 2) + (1. / ((1. - eps) ** 2)) -
                                                      (1. / (eps * (1. - eps)))) / E ** 3, 0)


def gryz_dCS(E, W, nsi, Ebi):
    """extFunctions.py:189-219."""
    eps = W / E
    epsB = Ebi / E
    with np.errstate(all='ignore'):
        return np.where((eps < 1.) & (epsB < 1.),
                        nsi * PI_EFOUR * eps * (1. + epsB) ** (-1.5) * (1. - eps) ** (epsB / (epsB + eps)) *
                        ((1. - epsB) + 4. * epsB * np.log(2.7 + ((1. - eps) / epsB) ** (0.5)) / (3. * eps)) / (W ** 3),
                        0)


def _orders_of_mag(max_size=1000000, num_iter=5):
    """probTables.py:54-58."""
    return [int(n) for n in np.geomspace(100, max_size, num=num_iter, dtype=int)]


def _percents(max_size, num_iter=10):
    """probTables.py:61-64."""
    return [int(n) for n in np.linspace(max_size / 10, max_size, num_iter, dtype=np.float32)]


class TableSearchExhausted(Exception):
    """The reference raises StopIteration here (probTables.py:241-244, 265-268)."""


class ProbTable:
    """probTables.py:70-535 restated without dask / memmap / masked arrays.

    ``table[iE, j]`` = CDF at ``Ws[j+1]`` for incident energy ``Es[iE]``; entries beyond the
    last valid step of a row are 0 (the on-disk convention, :500-507) and ``nvalid[iE]`` is the
    number of valid entries (the reference re-masks zeros on read, :533).
    """

    def __init__(self, kind, mat, E_range, tol_E, tol_W, Wc=None, shell=None):
        self.kind, self.mat, self.tol_E, self.tol_W = kind, mat, tol_E, tol_W
        self.Ef = mat.fermi_e
        if kind == 'Moller':                                               # :106-119
            self.shell = mat.name_val
            self.Wmin = Wc
            self.numEl = mat.n_val
            self.func_EW = lambda E, W: moller_dCS(E, W, self.numEl)
            self.Emin, self.Emax = E_range
        else:                                                              # :121-142
            self.shell = shell
            i = mat.name_s.index(shell)
            self.Wmin = _SPECIES[mat.species][6][i][1]                     # python number as in material.py
            self.numEl = _SPECIES[mat.species][6][i][2]
            self.func_EW = lambda E, W: gryz_dCS(E, W, self.numEl, self.Wmin)
            self.Emin, self.Emax = (self.Wmin, E_range[1]) if E_range[0] < self.Wmin else E_range
        self.Ws = self.Es = self.table = self.nvalid = None

    def Wmax(self, E):
        """probTables.py:17-41."""
        return (E - self.Ef) * 0.5

    # -- axis searches ---------------------------------------------------------------
    def _scalar_integrand(self, E):
        return lambda W: float(self.func_EW(E, W))

    def _check_intTol(self, series, ref, E):
        """probTables.py:165-188.  The reference evaluates the integrand one np.float32 scalar at a
        time; evaluating it on the float32 array is the same arithmetic."""
        y = np.asarray(self.func_EW(E, series))
        trap = np.trapezoid(y, x=series)
        return abs(trap - ref) / ref < self.tol_W

    def _check_CDFTol(self, Eseries, ref):
        """probTables.py:191-232 (only ``totalInt`` enters the decision)."""
        E = Eseries[-2]
        sel = self.Ws[self.Ws <= self.Wmax(E)]
        total = np.trapezoid(self.func_EW(E, sel), x=sel)
        return abs(total - ref) / ref < self.tol_E

    def find_Wseries(self, strict=True):
        """probTables.py:280-328."""
        from scipy.integrate import quad
        E = self.Emax
        W_max = self.Wmax(E)
        ref, err = quad(self._scalar_integrand(E), self.Wmin, W_max, limit=300, epsabs=1e-34)
        assert err < self.tol_W * ref

        def search(cands):
            for n in cands:
                s = np.geomspace(self.Wmin, W_max, num=n, dtype=np.float32)
                if self._check_intTol(s, ref, E):
                    return n, s
            if strict:
                raise TableSearchExhausted('W series, %s %s' % (self.kind, self.shell))
            return n, s
        n1, _ = search(_orders_of_mag())
        _, series = search(_percents(n1))
        return series

    def find_Eseries(self, strict=True):
        """probTables.py:332-382."""
        from scipy.integrate import quad
        Eref = self.Emax
        total, _ = quad(self._scalar_integrand(Eref), self.Wmin, self.Wmax(Eref), limit=300, epsabs=1e-34)

        def search(cands, maker):
            for n in cands:
                s = maker(n)
                if self._check_CDFTol(s, total):
                    return n, s
            if strict:
                raise TableSearchExhausted('E series, %s %s' % (self.kind, self.shell))
            return n, s
        n1, _ = search(_orders_of_mag(), lambda n: np.geomspace(self.Emin, self.Emax, num=n, dtype=np.float32))
        _, series = search(_percents(n1), lambda n: np.linspace(self.Emin, self.Emax, num=n, dtype=np.float32))
        return series

    def set_axes(self, strict=True):
        self.Ws = self.find_Wseries(strict)                                  # :385-397
        assert np.all(self.Ws < self.Emax)
        self.Es = self.find_Eseries(strict)                                  # :400-411
        return self

    # -- rows ------------------------------------------------------------------------
    def _rows_integrand(self, E32, W32):
        """The integrand as computeBlock evaluates it (probTables.py:439-448 calling
        extFunctions.py:178-182 / :213-218 with a float32 ndarray E mesh and a float32 MASKED W mesh).
        Under numpy >= 2 the masked-array operators wrap python scalars in 0-d float64 arrays, so the
        arithmetic is mixed: eps = W/E (and epsB, eps+epsB, their ratio, E**3, (1+epsB)**-1.5) in
        float32, everything downstream of the first scalar-on-masked operation in float64."""
        f8 = np.float64
        eps32 = W32 / E32
        eps = eps32.astype(f8)
        if self.kind == 'Moller':
            E3 = (E32 ** 3).astype(f8)                                        # ndarray float32 power
            val = self.numEl * PI_EFOUR * (1. / np.power(eps, 2.) + (1. / np.power(1. - eps, 2.)) -
                                           (1. / (eps * (1. - eps)))) / E3
            return np.where(eps32 < 1., val, 0.)
        epsB32 = (self.Wmin / E32).astype(np.float32)
        expo = (epsB32 / (epsB32 + eps32)).astype(f8)                         # float32 ratio
        pre = ((1. + epsB32) ** (-1.5)).astype(f8)                            # float32 power
        one_m_eps = 1. - eps
        with np.errstate(all='ignore'):
            val = (self.numEl * PI_EFOUR * eps) * pre * np.power(one_m_eps, expo) * \
                ((1. - epsB32).astype(f8) +
                 (4. * epsB32).astype(f8) * np.log(2.7 + np.power(one_m_eps / epsB32.astype(f8), 0.5)) / (3. * eps)) / \
                np.power(W32.astype(f8), 3.)
        return np.where((eps32 < 1.) & (epsB32 < 1.), val, 0.)

    def compute_rows(self, Es_block):
        """probTables.py:415-462: trapezoid steps masked where either end has W > Wmax(E) (:439-454),
        sequential cumsum (:457), divide by the row total (:460); float64 from the integrand on,
        cast to float32 when stored (mapToMemory :500-507)."""
        Es_block = np.asarray(Es_block, dtype=np.float32)
        Ws = self.Ws
        wmax = (Es_block - self.Ef) * 0.5                                    # float32 (weak python scalars)
        E_mesh, W_mesh = np.meshgrid(Es_block, Ws, indexing='ij')
        invalid = W_mesh > wmax[:, None]
        with np.errstate(all='ignore'):
            f = self._rows_integrand(E_mesh, W_mesh)
        d = W_mesh[:, 1:] - W_mesh[:, :-1]                                   # float32 diff of the W stack
        step = d * (f[:, 1:] + f[:, :-1]) / 2.0                              # np.trapz over the 2-stack, float64
        step_invalid = invalid[:, 1:] | invalid[:, :-1]
        step = np.where(step_invalid, 0., step)
        cum = np.cumsum(step, axis=1)
        with np.errstate(all='ignore'):
            cdf = (cum / cum[:, -1:]).astype(np.float32)
        cdf[step_invalid] = 0
        return cdf, (~step_invalid).sum(axis=1)

    def generate(self, chunk=100):
        """probTables.py:465-486 (chunking does not change values)."""
        rows, nv = [], []
        for i in range(0, len(self.Es), chunk):
            c, n = self.compute_rows(self.Es[i:i + chunk])
            rows.append(c)
            nv.append(n)
        self.table = np.concatenate(rows, axis=0)
        # readFromMemory (:533) masks every stored zero, so CDF entries that underflowed to 0 at the
        # start of a row are masked too; bisect then treats them as +inf.  Record the mask explicitly.
        self.masked = (self.table == 0)
        self.nvalid = np.concatenate(nv)
        return self


def gen_tables(inputPar, mat=None, strict=True, gryz='outermost'):
    """probTables.py:541-614.  ``tables['Gryz'][0]`` is the LAST (outermost) shell (:579).
    With ``gryz='outermost'`` only the table the hot path reads is built."""
    mat = mat or Material(inputPar['material'])
    Erange = (inputPar['Emin'], inputPar['E0'])
    t = {}
    t['Moller'] = ProbTable('Moller', mat, Erange, inputPar['tol_E'], inputPar['tol_W'],
                            Wc=inputPar['Wc']).set_axes(strict).generate()
    shells = mat.name_s[::-1]
    if gryz == 'outermost':
        shells = shells[:1]
    t['Gryz'] = [ProbTable('Gryzinski', mat, Erange, inputPar['tol_E'], inputPar['tol_W'],
                           shell=s).set_axes(strict).generate() for s in shells]
    if 'Mott' in inputPar.get('elastic', ''):                              # probTables.py:601-613
        t['Mott'] = MottTable(mat.Z, mat.species, inputPar.get('mott_path', '../tables/mott/')).read_ioffe().to_prob_table()
    return t


def pick_table(tab, E, u):
    """scattering.py:63-94.  bisect_left on Es (round up), bisect_left on the masked row
    (masked entries compare as 'not less' => +inf), W = Ws[j] (lower bin edge)."""
    iE = bisect.bisect_left(tab.Es, E)
    row = tab.table[iE]
    msk = tab.masked[iE]
    lo, hi = 0, len(row)
    while lo < hi:                       # bisect_left with masked-as-not-less semantics
        mid = (lo + hi) // 2
        if (not msk[mid]) and row[mid] < u:
            lo = mid + 1
        else:
            hi = mid
    return tab.Ws[lo], iE, lo


# --------------------------------------------------------------------------------------
# Philox4x32-10 counter-based RNG (shared definition with the CUDA path:
# key = (seed_lo, seed_hi), counter = (el_id_lo, el_id_hi, event_idx, stream))
# --------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85
_M32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _PH_M0 * c0
        p1 = _PH_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0 = (k0 + _PH_W0) & _M32
        k1 = (k1 + _PH_W1) & _M32
    return c0, c1, c2, c3


def u01(x):
    """32 random bits -> uniform in (0,1) that is EXACT in float32: ((x >> 9) + 0.5) * 2^-23
    (23 random bits plus the half-offset fill the 24-bit significand; never 0 or 1)."""
    return ((x >> 9) + 0.5) * (1.0 / 8388608.0)


def event_uniforms(seed, el_id, ev_idx):
    r = philox4x32_10((el_id & _M32, (el_id >> 32) & _M32, ev_idx & _M32, 0), (seed & _M32, (seed >> 32) & _M32))
    return tuple(u01(x) for x in r)


# --------------------------------------------------------------------------------------
# trajectories (singleScatter.py:10-41, 205-239; electron.py:44-113)
# --------------------------------------------------------------------------------------
class Electron:
    """electron.py:7-113 state + the rows it records (column contract fileTools.py:193-242)."""

    def __init__(self, E0, Emin, pos0, dir0):
        self.energy, self.Emin = E0, Emin
        self.xyz = np.array(pos0, float)
        self.dir = np.array(dir0, float)
        self.y_local = np.array([0., 1., 0.])
        self.outcome = None
        self.els = []      # rows: (outcome, final_E, dir(3), last_pos(3))
        self.scats = []    # rows: dict per event

    def save_outcomes(self):                                               # electron.py:105-113
        self.els.append((self.outcome, float(self.energy), self.dir.copy(), self.xyz.copy()))

    def update_xyz(self, pathl, zmax):                                     # electron.py:64-84
        new = self.xyz + pathl * self.dir
        if new[2] <= 0.:
            self.outcome = 'bks'
            self.save_outcomes()
        elif new[2] >= zmax:
            self.outcome = 'trsm'
            self.save_outcomes()
        else:
            self.xyz = new

    def update_energy(self, W):                                            # electron.py:44-59
        energy = self.energy - W
        if energy <= self.Emin:
            self.outcome = 'abs'
            self.save_outcomes()
        self.energy = energy


def initial_dir(s_tilt_deg):
    """multiScatter.py:97-100."""
    t = np.radians(s_tilt_deg)
    return np.array([-np.sin(t), 0., np.cos(t)])


def ds_event(el, mat, Wc, thickness, tables, us, gryz_per_shell=False, elastic='Ruth_vanilla'):
    """One iteration of the trajectory_DS loop body after the maxScatt check
    (singleScatter.py:25-41 + scattering.py:530-679).  ``us`` = (u_type, u_path, u_channel, u_phi);
    u_channel is unused for Quinn.  Returns the scats row."""
    E = el.energy
    mott = tables.get('Mott') if elastic == 'Mott' else None
    sig = sigmas_ds(E, mat, Wc, elastic, mott)
    mfp = mfp_from_sigma(sum(sig), mat.atnd)                               # scattering.py:519
    row = {'position': el.xyz.copy(), 'E': E}
    ch = pick_from_sigmas(sig, us[0])                                      # :530-549
    row['type'] = 'Mott' if (ch == 0 and elastic == 'Mott') else mat.channel_names[ch]
    row['channel'] = ch
    pathl = -mfp * math.log(us[1])                                         # :551-565
    row['pathl'] = pathl
    el.update_xyz(pathl, thickness)
    nsh = len(mat.Es)
    if ch == 0 and elastic == 'Mott':                                      # Mott :588-591
        W = 0.
        c2, _, _ = pick_mott_table(mott, E, us[2])
    elif ch == 0:                                                          # Ruth :575-577
        W = 0.
        c2 = rutherford_halfpol(E, mat.Z, us[2])
    elif ch == 1:                                                          # Moller :602-609
        W, _, _ = pick_table(tables['Moller'], E, us[2])
        W = float(W)                                                       # SURVEY Q9 (author-era float64)
        c2 = binary_coll(E, W, mat.fermi_e)
    elif ch < 2 + nsh:                                                     # Gryz :621-631 (table [0], quirk Q1)
        gi = 0
        if gryz_per_shell:
            gi = (nsh - 1) - (ch - 2)                                      # tables['Gryz'] is reversed name_s
        W, _, _ = pick_table(tables['Gryz'][gi], E, us[2])
        W = float(W)
        c2 = binary_coll(E, W, mat.fermi_e)
    else:                                                                  # Quinn :643-647
        W = mat.plasmon_e
        c2 = 1.
    halfPhi = math.pi * us[3]                                              # :676
    row.update(E_loss=W, pol_angle=c2, az_angle=halfPhi)
    el.update_energy(W)
    row['energy'] = el.energy
    el.dir, el.y_local = update_direction(el.dir, el.y_local, c2, halfPhi)
    row['direction'] = el.dir.copy()
    return row


def trajectory_ds(el, mat, Wc, maxScatt, thickness, tables, uniforms, gryz_per_shell=False,
                  elastic='Ruth_vanilla'):
    """singleScatter.py:10-41.  ``uniforms(ev_idx)`` -> 4 uniforms for that event."""
    n = 0
    while el.outcome not in OUTCOMES:
        n += 1
        if n > maxScatt:                                                   # :19-23 (dangling position dropped, Q6)
            el.outcome = 'long'
            el.save_outcomes()
            return
        el.scats.append(ds_event(el, mat, Wc, thickness, tables, uniforms(n - 1), gryz_per_shell, elastic))


def cont_event(el, mat, thickness, model, us, elastic='Ruth_vanilla'):
    """trajectory_cont_cl loop body (singleScatter.py:220-239 + scattering.py:802-840).
    ``us`` = (u_path, u_polar, u_phi)."""
    E = el.energy
    mfp = mfp_from_sigma(elastic_sigma(E, mat.Z, elastic, 'cont'), mat.atnd)   # scattering.py:781-792
    pathl = -mfp * math.log(us[0])                                         # :807
    row = {'pathl': pathl}
    el.update_xyz(pathl, thickness)
    W = pathl * stopping_power(model, E, mat)                              # :820
    row['E_loss'] = W
    el.update_energy(W)
    row['energy'] = el.energy
    c2 = rutherford_halfpol(E, mat.Z, us[1])                               # :829
    halfPhi = math.pi * us[2]                                              # :831
    row.update(type=elastic, channel=0, pol_angle=c2, az_angle=halfPhi)    # :834
    el.dir, el.y_local = update_direction(el.dir, el.y_local, c2, halfPhi)
    row['direction'] = el.dir.copy()
    return row


def trajectory_cont(el, mat, maxScatt, thickness, model, uniforms, elastic='Ruth_vanilla'):
    """singleScatter.py:205-239 (exits on abs/bks/trsm; 'long' returns)."""
    n = 0
    while el.outcome not in ('abs', 'bks', 'trsm'):
        n += 1
        if n > maxScatt:
            el.outcome = 'long'
            el.save_outcomes()
            return
        el.scats.append(cont_event(el, mat, thickness, model, uniforms(n - 1), elastic))


def run_electrons(inputPar, first_id, count, seed, tables=None, mat=None, gryz_per_shell=False):
    """multiTraj_DS / multiTraj_cont (multiScatter.py:66-116, 171-229) with the per-job
    ``random.seed(pid*time)`` (:94) replaced by Philox streams keyed on the electron id."""
    mat = mat or Material(inputPar['material'])
    d0 = initial_dir(inputPar['s_tilt'])
    elastic = inputPar.get('elastic', 'Ruth_vanilla')
    out = []
    for eid in range(first_id, first_id + count):
        el = Electron(inputPar['E0'], inputPar['Emin'], (0., 0., 0.), d0)
        uni = lambda ev, eid=eid: event_uniforms(seed, eid, ev)
        if inputPar['mode'] == 'DS':
            trajectory_ds(el, mat, inputPar['Wc'], inputPar['maxScatt'], inputPar['thickness'], tables, uni,
                          gryz_per_shell, elastic)
        else:
            u3 = lambda ev, uni=uni: uni(ev)[:3]
            trajectory_cont(el, mat, inputPar['maxScatt'], inputPar['thickness'], inputPar['Bethe_model'], u3,
                            elastic)
        out.append(el)
    return out
