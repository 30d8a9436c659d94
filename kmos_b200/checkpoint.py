"""The reference's restart file (<system_name>.reload), read and written per replica.

Format: base.save_system (kmos/fortran_src/base.mpy:517-578) -- '#' comments, then one labelled line per
array: kmc_time, walltime, kmc_step, nr_of_proc, volume, procstat, nr_of_sites, rates, lattice, and per process
`avail_sites i ...` (sites in list order) and `avail_sites_back i ...` (position of every site).  A file written
here loads into CPU kmos (base.reload_system, :365-510, skips labels it does not know) and vice versa.

Additions, all on lines the reference ignores: `integ_rates` (the reference forgets them, so a reloaded CPU run
restarts its coverage/TOF integrals; here they survive), `kmc_time_exact` (es22.15 drops the last bit of a
double) and an optional `#philox seed replica` comment.  `rates` are
written with the reference's 8 significant digits and therefore NOT read back: the batch keeps the rate constants
it was given, which is what makes a resumed trajectory bit-identical to an uninterrupted one.
"""
import numpy as np


def write_reload(path, state):
    """state: dict with kmc_time, kmc_step, procstat[P], nr_of_sites[P], rates[P], lattice[V], avail_sites[P][V][2],
    optional integ_rates[P], walltime, seed, replica."""
    P, V = len(state["procstat"]), len(state["lattice"])
    av = np.asarray(state["avail_sites"]).reshape(P, V, 2)
    with open(path, "w") as f:
        f.write("#Reload file written by kmos. Do not edit manually!\n")
        f.write("#Scalar variables\n")
        f.write(" kmc_time  %22.15E\n" % float(state["kmc_time"]))
        f.write(" walltime   %13.7E\n" % float(state.get("walltime", 0.0)))
        f.write(" kmc_step %22d\n" % int(state["kmc_step"]))
        f.write(" nr_of_proc %11d\n" % P)
        f.write(" volume %11d\n" % V)
        if "seed" in state:
            f.write("#philox %d %d\n" % (int(state["seed"]), int(state.get("replica", 0))))
        f.write("#Vector variables\n")
        f.write("procstat " + "".join("%21d" % int(x) for x in state["procstat"]) + "\n")
        f.write("nr_of_sites " + "".join("%9d" % int(x) for x in state["nr_of_sites"]) + "\n")
        f.write("rates " + "".join("%14.7E" % float(x) for x in state["rates"]) + "\n")
        f.write("kmc_time_exact %s\n" % float(state["kmc_time"]).hex())
        if state.get("integ_rates") is not None:
            f.write("integ_rates " + " ".join(float(x).hex() for x in state["integ_rates"]) + "\n")
        f.write("lattice " + "".join("%9d" % int(x) for x in state["lattice"]) + "\n")
        for i in range(P):
            f.write("avail_sites " + "%9d" % (i + 1) + "".join("%9d" % int(x) for x in av[i, :, 0]) + "\n")
        for i in range(P):
            f.write("avail_sites_back " + "%9d" % (i + 1) + "".join("%9d" % int(x) for x in av[i, :, 1]) + "\n")


def read_reload(path):
    """-> dict like write_reload's argument (arrays as numpy)."""
    out, rows, back = {}, {}, {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith("#philox"):
                _tag, seed, rep = line.split()
                out["seed"], out["replica"] = int(seed), int(rep)
                continue
            if not line or line.startswith("#"):
                continue
            label, _, rest = line.partition(" ")
            vals = rest.split()
            if label in ("kmc_time", "walltime"):
                out[label] = float(vals[0])
            elif label in ("kmc_step", "nr_of_proc", "volume"):
                out[label] = int(vals[0])
            elif label == "procstat":
                out[label] = np.array(vals, dtype=np.int64)
            elif label in ("nr_of_sites", "lattice"):
                out[label] = np.array(vals, dtype=np.int32)
            elif label == "rates":
                out[label] = np.array(vals, dtype=np.float64)
            elif label == "kmc_time_exact":
                out[label] = float.fromhex(vals[0])
            elif label == "integ_rates":
                out[label] = np.array([float.fromhex(v) for v in vals])
            elif label == "avail_sites":
                rows[int(vals[0])] = np.array(vals[1:], dtype=np.int32)
            elif label == "avail_sites_back":
                back[int(vals[0])] = np.array(vals[1:], dtype=np.int32)
    if "kmc_time_exact" in out:  # es22.15 drops the last bit of a double
        out["kmc_time"] = out.pop("kmc_time_exact")
    P, V = out["nr_of_proc"], out["volume"]
    av = np.zeros((P, V, 2), dtype=np.int32)
    for i in range(1, P + 1):
        av[i - 1, :, 0] = rows[i]
        av[i - 1, :, 1] = back[i]
    out["avail_sites"] = av
    return out
