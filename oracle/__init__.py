"""TEST INFRASTRUCTURE: CPU oracle (see oracle.cpp header).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
