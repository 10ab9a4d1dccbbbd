import numpy as np


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    m = np.abs(b) > 0
    if not m.any():
        return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def tau_relerr(tau, tau_ref, last):
    """Compare optical-depth columns only down to `last` (the reference leaves the rest at 0)."""
    worst = 0.0
    for w in range(tau_ref.shape[0]):
        L = int(last[w])
        worst = max(worst, relerr(tau[w, :L + 1], tau_ref[w, :L + 1]))
    return worst


def apply_setters(obj, setters):
    if "radius" in setters:
        obj.set_radius(setters["radius"])
    if "cloudtop" in setters:
        obj.set_cloudtop(setters["cloudtop"])
    if "scattering" in setters or "scatflag" in setters:
        obj.set_scattering(setters.get("scatflag", 1), setters.get("scattering", 0.0))


def parse_dump(path):
    """One of transit's `savefiles` text dumps (tau.c:360-518) -> (keys, rows): every record is a
    'wavenumber: x' / 'radius: x' line followed by one line of values.  tau.dat parsed this way is
    what code/cf.py:68-96 (readTauDat) extracts."""
    keys, rows = [], []
    with open(path) as f:
        lines = f.readlines()
    i = 0
    while i < len(lines):
        s = lines[i].strip()
        if s.startswith("wavenumber:") or s.startswith("radius:"):
            keys.append(float(s.split()[1]))
            rows.append([float(v) for v in lines[i + 1].split()])
            i += 2
        else:
            i += 1
    return np.array(keys), np.array(rows)


DUMPS = ("tau.dat", "CIA.dat", "mol_extion.dat", "total_extion.dat", "cloud_extion.dat",
         "scatt_extion.dat")


class TorchComm:
    """All-gather through torch.distributed (backend gloo on CPU, nccl on GPUs)."""

    def __init__(self, dist, device="cpu"):
        self.dist, self.device = dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allgather(self, local, counts):
        import torch
        width = local.shape[1]
        pad = max(counts)
        buf = torch.zeros((pad, width), dtype=torch.float64, device=self.device)
        buf[:local.shape[0]] = torch.as_tensor(local, dtype=torch.float64)
        outs = [torch.zeros_like(buf) for _ in range(self.world)]
        self.dist.all_gather(outs, buf)
        return np.concatenate([o[:n].cpu().numpy() for o, n in zip(outs, counts)], axis=0)
