chemical_symbols = ['X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg', 'Al', 'Si',
                    'P', 'S', 'Cl', 'Ar']
atomic_numbers = {s: i for i, s in enumerate(chemical_symbols)}
# IUPAC 2016 abridged (what ase.data.atomic_masses holds since ASE 3.13)
atomic_masses = [1.0, 1.008, 4.002602, 6.94, 9.0121831, 10.81, 12.011, 14.007, 15.999, 18.998403163,
                 20.1797, 22.98976928, 24.305, 26.9815385, 28.085, 30.973761998, 32.06, 35.45, 39.948]
# pre-3.13 ASE values, kept for the golden-replay search in tests
atomic_masses_legacy = [0.0, 1.00794, 4.002602, 6.941, 9.012182, 10.811, 12.0107, 14.0067, 15.9994,
                        18.9984032, 20.1797, 22.98976928, 24.3050, 26.9815386, 28.0855, 30.973762,
                        32.065, 35.453, 39.948]
