import numpy as np


class Atoms(object):
    def __init__(self, symbols=None, positions=None, *args, **kwargs):
        self.symbols = symbols
        self.positions = positions

    def get_masses(self):
        from ase.symbols import string2symbols
        from ase.data import atomic_masses, atomic_numbers
        if isinstance(self.symbols, str):
            return np.array([atomic_masses[atomic_numbers[s]] for s in string2symbols(self.symbols)])
        return np.array([1.0])

    def get_chemical_formula(self, mode="hill"):
        return self.symbols if isinstance(self.symbols, str) else "X"

    def get_name(self):
        return self.get_chemical_formula()

    def __len__(self):
        return len(self.get_masses())

    def __eq__(self, other):
        return self is other

    def __hash__(self):
        return id(self)
