"""CPU oracle for the laser-polio per-tick agent update.  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import anything from this package."""
