"""CPU arm of bench.py: loader for the UNMODIFIED reference installed under baseline/_ref (git-ignored; see
DESIGN.md "Measurement")."""
