"""TEST STUB (see ase/__init__.py): the one constant the calculators use (CODATA 2014, as ase.units)."""
Hartree = 27.211386024367243
