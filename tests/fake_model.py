"""A stand-in for kmos_b200.model.KMC_Model with the methods ModelRunner uses: rows are a pure function of the
parameter point and the Philox key, so sharded and unsharded scans can be compared on a machine without a GPU."""
import numpy as np


class FakeModel(object):
    def __init__(self, model, size=20, n_replicas=1, parameters=None, device=0, seeds=None, replica_ids=None,
                 gpu_ids=None):
        self.params, self.seeds, self.device = parameters, np.asarray(seeds, dtype=np.uint64), device
        if gpu_ids is not None:  # a fleet numbers its replicas 0..R-1 itself
            assert replica_ids is None and len(gpu_ids) > 0
            replica_ids = np.arange(n_replicas)
        assert np.array_equal(np.asarray(replica_ids, dtype=np.uint64) + np.uint64(7), self.seeds)  # global ids
        self.steps = 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def do_steps(self, n):
        self.steps += int(n)

    def get_atoms_all(self):
        return None

    def get_std_sampled_data_all(self, samples, sample_steps, tof_method="integ"):
        rows = []
        for p, s in zip(self.params, self.seeds):
            rows.append([p["T"], p["p_COgas"], float(s) * 1e-3, float(self.steps + sample_steps)])
        return np.asarray(rows)

    def get_std_header(self):
        return "#T p_COgas tof kmc_steps\n"

    def get_parameters(self, replica=0):
        return dict(self.params[replica])
