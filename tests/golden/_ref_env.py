"""Build an importable copy of the *reference* jitr in a temp dir (generation-time only).

The reference lives read-only at /root/reference and cannot be imported in place:
it needs a generated ``__version__.py`` and the ``periodictable`` package (absent,
used for element-symbol strings only).  This helper copies ``src/jitr`` to a temp
dir, symlinks the data directory, writes the two shims and prepends the temp
paths to ``sys.path``.  Only ``make_golden.py`` (and ad-hoc validation here in the
build container) uses it -- nothing under ``tests/`` that runs on the GPU box
may import this module, because /root/reference does not exist there.
"""
import os
import shutil
import sys
import tempfile

REFERENCE = "/root/reference"


def setup(tmp=None):
    if tmp is None:
        tmp = os.path.join(tempfile.gettempdir(), "jitr_ref_env")
    src = os.path.join(tmp, "src")
    stubs = os.path.join(tmp, "stubs")
    if not os.path.exists(os.path.join(src, "jitr", "__version__.py")):
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(src)
        os.makedirs(os.path.join(stubs, "periodictable"))
        shutil.copytree(os.path.join(REFERENCE, "src", "jitr"), os.path.join(src, "jitr"))
        os.symlink(os.path.join(REFERENCE, "src", "data"), os.path.join(src, "data"))
        with open(os.path.join(src, "jitr", "__version__.py"), "w") as f:
            f.write('__version__ = "0+oracle"\n')
        with open(os.path.join(stubs, "periodictable", "__init__.py"), "w") as f:
            f.write(
                "_SYM = ('n H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn '\n"
                "        'Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr '\n"
                "        'Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr '\n"
                "        'Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc '\n"
                "        'Lv Ts Og').split()\n"
                "class _El:\n"
                "    def __init__(self, z):\n        self.number = z\n        self.symbol = _SYM[z]\n"
                "    def __str__(self):\n        return self.symbol\n"
                "    __repr__ = __str__\n"
                "class _Els:\n"
                "    def __getitem__(self, z):\n        return _El(int(z))\n"
                "elements = _Els()\n"
            )
    for p in (stubs, src):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    return tmp
