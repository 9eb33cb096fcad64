"""Host mirror of the reference's Mott table object (MC/tools/extFunctions.py:21-153) and of its sampling function
(MC/tools/scattering.py:96-122).

`elastic: Mott` makes genTables (probTables.py:601-613) read `../tables/mott/ioffe/<species>.ioffe`, a table of
numerical elastic cross-sections downloaded from ioffe.ru/ES/Elastic (not shipped with the reference), and hand the
object to the DS event loop, which takes the elastic cross-section sigma(E) and the polar angle from it.  Here the
object keeps the reference's attribute and method names; `upload(engine)` puts it on the device
(`mcd_mott_upload`), where the trajectory kernel does the two bisections per elastic event.

File layout the parser relies on (extFunctions.py:56-94): line 3 ends with Z, lines 6 / 7 start with the number of
angles / energies, line 13 is `theta [grad]= t0 t1 ...`, rows `E | dCS(t0) ... dCS(tn) | sigma` run from line 15 to
the last line but one, descending in energy.
"""
import bisect
import os
import random

import numpy as np


class mottTable:
    """sigma(E) in cm^2 and the cumulative distribution of the polar angle on the table's (E, theta) grid."""

    def __init__(self, Z, name, path='../tables/mott/'):
        self.Z, self.name = Z, name
        self.path = path
        self.target = os.path.join(self.path, name + '.table')
        self.ioffeFile = os.path.join(self.path + 'ioffe/' + name + '.ioffe')
        self.probTable = None
        self.thetas = None
        if not os.path.exists(self.ioffeFile):
            raise FileNotFoundError('Data for %s was not found (%s): Mott tables come from ioffe.ru/ES/Elastic and '
                                    'are not shipped' % (name, self.ioffeFile))
        head = self._lines()[:7]
        if int(head[2].split()[-1]) != self.Z:
            raise ValueError('Trying to read file %s but expecting Z:%s' % (self.ioffeFile, self.Z))
        self.sizeTheta, self.sizeE = int(head[5].split()[0]), int(head[6].split()[0])
        self.Es = np.empty(self.sizeE)
        self.sigmas = np.empty(self.sizeE)
        self.dCS = np.empty([self.sizeE, self.sizeTheta])

    def _lines(self):
        with open(self.ioffeFile, 'r') as f:
            return f.readlines()

    def readIoffe(self):
        """Parse angles, energies, dCS rows and total cross-sections; flip everything to ascending energy."""
        lines = self._lines()
        self.thetas = np.array(lines[12].strip('theta [grad]=').split(), dtype=float)
        body = np.array([ln.replace('|', ' ').split() for ln in lines[14:-1]], dtype=float)
        if body.shape != (self.sizeE, self.sizeTheta + 2) or len(self.thetas) != self.sizeTheta:
            raise ValueError('%s: %s rows of %s numbers for %d energies x %d angles'
                             % (self.ioffeFile, body.shape[0], body.shape[1:], self.sizeE, self.sizeTheta))
        self.Es = body[::-1, 0].copy()
        self.sigmas = body[::-1, -1].copy()
        self.dCS = body[::-1, 1:-1].copy()
        return self

    def toProbTable(self):
        """CDF over the angle grid: running sum of dCS divided by the row sum (no solid-angle weight, :99-109)."""
        run = np.cumsum(self.dCS, axis=1)
        self.probTable = run / np.sum(self.dCS, axis=1)[:, None]
        return self

    def mapToMemory(self):
        """Store the CDF as a float32 memmap next to the ioffe directory (:112-128)."""
        m = np.memmap(self.target, dtype='float32', mode='w+', shape=(self.sizeE, self.sizeTheta))
        m[:] = self.probTable[:]
        del m
        print('Table written to target:', self.target)

    def readFromMemory(self):
        """Load the float32 CDF memmap (:131-145); the axes and sigmas still come from the ioffe file."""
        if self.thetas is None:
            self.readIoffe()
        if os.path.exists(self.target):
            self.probTable = np.memmap(self.target, dtype='float32', mode='r', shape=(self.sizeE, self.sizeTheta))
            print('read Mott table from memory for %s\n' % self.name)
        else:
            self.generate()
        return self

    def generate(self):
        self.readIoffe()
        self.toProbTable()
        self.mapToMemory()
        return self

    def upload(self, engine):
        if self.probTable is None:
            self.readIoffe().toProbTable()
        engine.mott_upload(self.Es, self.sigmas, np.asarray(self.probTable, dtype=np.float64), self.thetas)
        return self


def pickMottTable(table, energy):
    """scattering.py:96-122 on the host (one draw from `random`): cos^2(theta/2) of a Mott event at `energy`."""
    row = table.probTable[bisect.bisect_left(table.Es, energy)]
    theta = table.thetas[bisect.bisect_left(row, random.random())]
    return (np.cos(np.radians(theta * 0.5))) ** 2
