"""kmos_b200 -- B200-native batched engine for the kmos kMC step loop.

Scope: the per-step event cycle of kmos' generated base/lattice/proclist Fortran modules
(update_accum_rate -> update_clocks -> update_integ_rate -> determine_procsite -> run_proc_nr),
for thousands of independent replicas at once.  See DESIGN.md.
"""
__version__ = "0.1.0"
