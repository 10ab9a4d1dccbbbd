"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of liboracle.so (oracle/transit_oracle.c) plus small, independent
Python/numpy readers of the reference's input formats, so that the oracle shares no parsing
code with the product library.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

The band-integration oracle (`resample`, `bandintegrate`) restates code/wine.py:127-199 and
code/BARTfunc.py:386-396 in numpy.
"""
import ctypes as C
import os
import struct
import subprocess
import numpy as np

_trapz = getattr(np, "trapezoid", None) or np.trapz

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "liboracle.so")
REFLIB = os.path.join(HERE, "_ref", "libtransit_ref.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIBPATH) or \
            os.path.getmtime(LIBPATH) < os.path.getmtime(os.path.join(HERE, "transit_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if os.path.isdir("/root/reference/modules/transit/transit/src") and \
            (force or not os.path.exists(REFLIB)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class OrcConfig(C.Structure):
    _fields_ = [
        ("nlayer", C.c_int), ("nspec", C.c_int), ("nwave", C.c_int),
        ("press", dp), ("pfct", C.c_double), ("rfct", C.c_double),
        ("mass", dp), ("pol", dp), ("wn", dp),
        ("gsurf", C.c_double), ("p0", C.c_double), ("r0", C.c_double),
        ("ntemp", C.c_int), ("ngmol", C.c_int),
        ("gtemp", dp), ("gmol_spec", ip), ("grid", dp),
        ("ncia", C.c_int), ("cia_nwn", ip), ("cia_nt", ip),
        ("cia_wn", C.POINTER(dp)), ("cia_t", C.POINTER(dp)), ("cia_tab", C.POINTER(dp)),
        ("cia_nspec", ip), ("cia_spec", ip),
        ("toomuch", C.c_double), ("nangle", C.c_int), ("angles_deg", dp),
        ("starrad_cm", C.c_double), ("transparent", C.c_int),
        ("cloud_flag", C.c_int), ("cloudext", C.c_double), ("cloudtop", C.c_double),
        ("cloudbot", C.c_double),
        ("scat_flag", C.c_int), ("scat_logext", C.c_double),
    ]


class OrcInter(C.Structure):
    _fields_ = [("radius", dp), ("temp", dp), ("mm", dp), ("dens", dp), ("ext", dp),
                ("cia", dp), ("tau", dp), ("last", C.POINTER(C.c_long)), ("intens", dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIBPATH)
        _lib.orc_forward.argtypes = [C.POINTER(OrcConfig), C.c_int, dp, dp, C.POINTER(OrcInter)]
        _lib.orc_forward.restype = C.c_int
        _lib.orc_simps_path.restype = C.c_double
        _lib.orc_simps_path.argtypes = [dp, dp, C.c_int]
        _lib.orc_totaltau1.restype = C.c_double
        _lib.orc_totaltau1.argtypes = [C.c_double, dp, dp, C.c_long]
        _lib.orc_modulation1.restype = C.c_double
        _lib.orc_modulation1.argtypes = [dp, C.c_long, C.c_double, dp, C.c_long, C.c_double,
                                         C.c_double, C.c_int]
        _lib.orc_voigtn.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(C.c_float), C.c_int]
        _lib.orc_profile_halfsize.restype = C.c_long
        _lib.orc_profile_halfsize.argtypes = [C.c_double, C.c_double, C.c_double, C.c_float,
                                              C.c_long]
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


# ----------------------------------------------------------------------------------------
# Independent readers of the reference's input formats
def read_cfg(path):
    """`key value` lines, '#' comments (procopt.c:649-704).  Returns the raw dict."""
    out = {}
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line or line.startswith("#"):
                continue
            parts = line.split(None, 1)
            out[parts[0]] = parts[1].strip() if len(parts) > 1 else ""
    return out


def read_molecules(path):
    tab = {}
    with open(path) as f:
        for line in f:
            if not line.strip() or line.lstrip().startswith("#"):
                continue
            p = line.split()
            tab[p[1]] = dict(id=int(p[0]), mass=float(p[2]), radius=float(p[3]) / 2.0 * 1e-8,
                             pol=float(p[5]))
    return tab


def read_atm(path):
    """TEA-format atmosphere file (readatm.c:255-620)."""
    species, rows = None, []
    rfct = pfct = tfct = 1.0
    with open(path) as f:
        lines = f.readlines()
    i = 0
    while i < len(lines):
        s = lines[i].strip()
        if s.startswith("#SPECIES"):
            species = lines[i + 1].split()
            i += 2
            continue
        if s.startswith("ur"):
            rfct = float(s[2:])
        elif s.startswith("up"):
            pfct = float(s[2:])
        elif s.startswith("ut"):
            tfct = float(s[2:])
        elif s and not s.startswith("#") and not s.startswith("q") and species is not None \
                and (s[0].isdigit() or s[0] in "+-."):
            rows.append([float(v) for v in s.split()])
        i += 1
    a = np.array(rows)
    return dict(species=species, radius=a[:, 0], press=a[:, 1], temp=a[:, 2], q=a[:, 3:],
                rfct=rfct, pfct=pfct, tfct=tfct)


def read_cia(path):
    species, temps, rows = None, None, []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            if s[0] == "i":
                species = s.split()[1:]
            elif s[0] == "t":
                temps = np.array([float(v.rstrip("kK")) for v in s.split()[1:]])
            else:
                rows.append([float(v) for v in s.split()])
    a = np.array(rows)
    return dict(species=species, temps=temps, wn=np.ascontiguousarray(a[:, 0]),
                tab=np.ascontiguousarray(a[:, 1:]))


def read_opacity(path):
    with open(path, "rb") as f:
        nmol, nT, nlayer, nw = struct.unpack("4l", f.read(32))
        molids = np.frombuffer(f.read(4 * nmol), dtype=np.int32).copy()
        temps = np.frombuffer(f.read(8 * nT), dtype=np.float64).copy()
        press = np.frombuffer(f.read(8 * nlayer), dtype=np.float64).copy()
        wn = np.frombuffer(f.read(8 * nw), dtype=np.float64).copy()
        off = f.tell()
    o = np.fromfile(path, dtype=np.float64, offset=off).reshape(nlayer, nT, nmol, nw)
    return dict(molids=molids, temps=temps, press=press, wn=wn, o=o)


def wn_grid(lo, hi, d):
    n = int(((1.0 + 1e-8) * hi - lo) / d + 1)
    return lo + np.arange(n) * d


class Oracle:
    """One configuration (the state transit_init leaves behind) for orc_forward."""

    def __init__(self, cfgpath):
        cfg = read_cfg(cfgpath)
        self.cfg = cfg
        key = lambda k, d=None: next((v for kk, v in cfg.items() if k.startswith(kk) or
                                      kk.startswith(k)), d) if k not in cfg else cfg[k]
        self.eclipse = cfg.get("solution", "eclipse") == "eclipse"
        atm = read_atm(cfg["atm"])
        mols = read_molecules(cfg["molfile"])
        self.atm = atm
        self.species = atm["species"]
        ns = len(self.species)
        self.mass = np.array([mols[s]["mass"] for s in self.species])
        self.pol = np.array([mols[s]["pol"] for s in self.species])
        self.radius_cm = np.array([mols[s]["radius"] for s in self.species])
        self.ids = [mols[s]["id"] for s in self.species]
        self.press = np.ascontiguousarray(atm["press"])
        if "wnlow" in cfg:
            lo, hi = float(cfg["wnlow"]), float(cfg["wnhigh"])
        else:
            fct = float(cfg.get("wlfct", 1e-4))
            lo, hi = 1.0 / (float(cfg["wlhigh"]) * fct), 1.0 / (float(cfg["wllow"]) * fct)
        self.wn = wn_grid(lo, hi, float(cfg["wndelt"]))
        c = OrcConfig()
        c.nlayer, c.nspec, c.nwave = len(self.press), ns, len(self.wn)
        c.press, c.pfct, c.rfct = _d(self.press), atm["pfct"], atm["rfct"]
        c.mass, c.pol, c.wn = _d(self.mass), _d(self.pol), _d(self.wn)
        c.gsurf = float(cfg["gsurf"])
        c.p0 = float(cfg["refpress"])
        c.r0 = float(cfg["refradius"])
        op = read_opacity(cfg["opacityfile"])
        self.op = op
        self.grid = np.ascontiguousarray(op["o"])
        self.gtemp = op["temps"]
        self.gmol_spec = np.array([self.ids.index(int(m)) for m in op["molids"]], dtype=np.int32)
        c.ntemp, c.ngmol = len(self.gtemp), len(self.gmol_spec)
        c.gtemp, c.gmol_spec, c.grid = _d(self.gtemp), self.gmol_spec.ctypes.data_as(ip), \
            _d(self.grid)
        files = [f for f in cfg.get("csfile", "").split(",") if f]
        self.cia = [read_cia(f) for f in files]
        n = len(self.cia)
        c.ncia = n
        self._keep = []
        if n:
            self.cia_nwn = np.array([len(t["wn"]) for t in self.cia], dtype=np.int32)
            self.cia_nt = np.array([len(t["temps"]) for t in self.cia], dtype=np.int32)
            self.cia_nspec = np.array([len(t["species"]) for t in self.cia], dtype=np.int32)
            spec = np.zeros(2 * n, dtype=np.int32)
            for i, t in enumerate(self.cia):
                for k, s in enumerate(t["species"]):
                    spec[2 * i + k] = self.species.index(s)
            self.cia_spec = spec
            wnp = (dp * n)(*[_d(t["wn"]) for t in self.cia])
            tp = (dp * n)(*[_d(t["temps"]) for t in self.cia])
            tabp = (dp * n)(*[_d(t["tab"]) for t in self.cia])
            self._keep += [wnp, tp, tabp]
            c.cia_nwn, c.cia_nt = self.cia_nwn.ctypes.data_as(ip), self.cia_nt.ctypes.data_as(ip)
            c.cia_wn, c.cia_t, c.cia_tab = wnp, tp, tabp
            c.cia_nspec, c.cia_spec = self.cia_nspec.ctypes.data_as(ip), spec.ctypes.data_as(ip)
        c.toomuch = float(cfg.get("toomuch", 20))
        self.angles = np.array([float(v) for v in cfg.get("raygrid", "0 20 40 60 80").split()])
        c.nangle, c.angles_deg = len(self.angles), _d(self.angles)
        c.starrad_cm = float(cfg.get("starrad", 1.125)) * 6.957e10
        c.transparent = 0
        c.cloud_flag, c.cloudext, c.cloudtop, c.cloudbot = 0, 0.0, 0.0, 0.0
        if "cloudtop" in cfg:
            self.set_cloudtop(float(cfg["cloudtop"]), c)
        c.scat_flag, c.scat_logext = 0, 0.0
        if "scattering" in cfg:
            if cfg["scattering"] == "polar":
                c.scat_flag = 2
            else:
                c.scat_flag, c.scat_logext = 1, float(cfg["scattering"])
        self.c = c

    # transit.c:98-115
    def set_radius(self, r):
        self.c.r0 = r

    def set_cloudtop(self, top, c=None):
        c = c or self.c
        c.cloud_flag, c.cloudext, c.cloudtop, c.cloudbot = 1, 100.0, top, top + 10

    def set_scattering(self, flag, logext):
        self.c.scat_flag, self.c.scat_logext = flag, logext

    def run(self, model, inter=False):
        model = np.ascontiguousarray(model, dtype=np.float64)
        nl, nw, ns = self.c.nlayer, self.c.nwave, self.c.nspec
        spec = np.zeros(nw)
        if not inter:
            lib().orc_forward(C.byref(self.c), int(self.eclipse), _d(model), _d(spec), None)
            return spec
        it = OrcInter()
        out = dict(radius=np.zeros(nl), temp=np.zeros(nl), mm=np.zeros(nl),
                   density=np.zeros((ns, nl)), ext=np.zeros((nl, nw)), cia=np.zeros((nw, nl)),
                   tau=np.zeros((nw, nl)), last=np.zeros(nw, dtype=np.int64))
        it.radius, it.temp, it.mm, it.dens = _d(out["radius"]), _d(out["temp"]), \
            _d(out["mm"]), _d(out["density"])
        it.ext, it.cia, it.tau = _d(out["ext"]), _d(out["cia"]), _d(out["tau"])
        it.last = out["last"].ctypes.data_as(C.POINTER(C.c_long))
        if self.eclipse:
            out["intens"] = np.zeros((self.c.nangle, nw))
            it.intens = _d(out["intens"])
        lib().orc_forward(C.byref(self.c), int(self.eclipse), _d(model), _d(spec), C.byref(it))
        out["spectrum"] = spec
        return out

    def run_batch(self, models):
        return np.stack([self.run(m) for m in np.atleast_2d(models)])


# ----------------------------------------------------------------------------------------
# Stage (c): band integration (code/wine.py:16-66,127-199; code/BARTfunc.py:386-396)
def readfilter(path):
    with open(path) as f:
        lines = f.readlines()
    while lines[0].startswith("#") or not lines[0].strip():
        lines.pop(0)
    n = len(lines)
    wl = np.zeros(n)
    tr = np.zeros(n)
    for i in range(n):
        a, b = lines[i].strip().split()[0:2]
        wl[n - 1 - i], tr[n - 1 - i] = float(a), float(b)
    return 1.0 / (wl * 1e-4), tr


def resample(specwn, filterwn, filtertr, starwn, starfl):
    idx = np.where((specwn < filterwn[-1]) & (filterwn[0] < specwn))[0]
    ifilter = np.interp(specwn[idx], filterwn, filtertr)
    istarfl = np.interp(specwn[idx], starwn, starfl)
    nifilter = ifilter / _trapz(ifilter, specwn[idx])
    return nifilter, istarfl, idx


def bandintegrate(spectrum, specwn, nifilter, idx):
    return _trapz(spectrum * nifilter, specwn[idx])


def bandflux(spectrum, specwn, filters, star=None, rprs=None):
    """filters: list of (nifilter, istarfl, idx).  Eclipse when `star` is truthy."""
    out = np.zeros(len(filters))
    for i, (nif, istar, idx) in enumerate(filters):
        if star:
            out[i] = bandintegrate(spectrum[idx] / istar * rprs * rprs, specwn, nif, idx)
        else:
            out[i] = bandintegrate(spectrum[idx], specwn, nif, idx)
    return out
