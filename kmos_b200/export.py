"""Exporter hook: emit the device rule tables next to the Fortran that ``kmos export`` writes.

The reference's ``kmos.io.export_source(project, export_dir, code_generator)`` (kmos/io/__init__.py:3884-3974)
writes ``base/lattice/proclist[...].f90`` into ``export_dir``; this module adds ``model_tables.json`` (the
rule-table IR) and ``model_tables.bin`` (the int32 image ``kmos_b200_model_create`` consumes) to the same
directory, derived from exactly those generated sources so that statement order -- and with it the order of
``avail_sites`` -- is the one the Fortran build would have.  INTEGRATION.md shows the two-line call a
maintainer adds at io/__init__.py:3958-3973.

    python -m kmos_b200.export <export_dir> [--backend local_smart|lat_int|otf]
"""
import json
import os
import sys

from . import fortran_ir, tables


def export_tables(export_dir, backend=None, project=None):
    """Parse the generated Fortran in `export_dir` and write model_tables.{json,bin} there."""
    ir = fortran_ir.parse_export_dir(export_dir, backend)
    if project is not None:
        ir.update(project_meta(project))
    ir.update(lattice_geometry(export_dir))
    blob, info = tables.build_blob(ir)
    with open(os.path.join(export_dir, "model_tables.json"), "w") as f:
        json.dump(ir, f, separators=(",", ":"), sort_keys=True)
    blob.tofile(os.path.join(export_dir, "model_tables.bin"))
    # the model's proclist as specialised CUDA, next to the proclist.f90 it mirrors (local_smart models the
    # generator takes; the others run on the table interpreter / the warp kernels)
    if ir["backend"] == "local_smart":
        from . import codegen, devtables
        try:
            path, _inf = codegen.write_source(ir, export_dir, blob)
            info["proclist_cu"] = path
        except devtables.Unsupported as e:
            info["proclist_cu"] = None
            info["proclist_cu_declined"] = str(e)
    return ir, blob, info


def lattice_geometry(export_dir):
    """unit_cell_size / site_positions from the generated lattice.f90 (kmos/io/__init__.py:650-700): only the
    front-end uses them (KMC_Model.cell_size, get_atoms), the step loop does not."""
    import re
    path = os.path.join(export_dir, "lattice.f90")
    if not os.path.exists(path):
        return {}
    cell = [[0.0] * 3 for _ in range(3)]
    pos = {}
    with open(path) as f:
        for line in f:
            m = re.match(r"\s*unit_cell_size\((\d), (\d)\) = ([-+.\deEdD]+)", line)
            if m:
                cell[int(m.group(1)) - 1][int(m.group(2)) - 1] = float(m.group(3).lower().replace("d", "e"))
            m = re.match(r"\s*site_positions\((\d+),:\) = \(/(.*)/\)", line)
            if m:
                pos[int(m.group(1))] = [float(x.lower().replace("d", "e")) for x in m.group(2).split(",")]
    return {"unit_cell_size": cell, "site_positions": [pos[k] for k in sorted(pos)]}


def project_meta(pt):
    """Rate expressions, parameters and TOF counters of a kmos Project (inputs of the hot path)."""
    params = {p.name: {"value": p.value, "adjustable": bool(p.adjustable), "min": p.min, "max": p.max,
                       "scale": p.scale} for p in pt.get_parameters()}
    procs = [{"name": proc.name, "rate_constant": proc.rate_constant, "otf_rate": getattr(proc, "otf_rate", None),
              "enabled": bool(proc.enabled), "tof_count": proc.tof_count if proc.tof_count else None}
             for proc in pt.get_processes()]
    return {"parameters": params, "process_defs": procs}


def export_source(project, export_dir=None, code_generator="local_smart", **kw):
    """Drop-in for kmos.io.export_source: the reference's export, then the device tables."""
    import kmos.io
    kmos.io.export_source(project, export_dir, code_generator=code_generator, **kw)
    return export_tables(export_dir, code_generator, project)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    backend = None
    for i, a in enumerate(sys.argv):
        if a == "--backend":
            backend = sys.argv[i + 1]
            args = [x for x in args if x != backend]
    ir, blob, info = export_tables(args[0], backend)
    print("%s: %d processes, backend %s, %d table words; shared-memory kernel: %s" % (
        args[0], len(ir["procs"]), ir["backend"], blob.size,
        "yes" if info["device"]["supported"] else "no (%s)" % info["device"].get("reason")))
