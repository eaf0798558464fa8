"""TEST INFRASTRUCTURE: CPU oracle of the hot path (recon_oracle.c) and runners for
the compiled, unmodified reference (oracle/_ref).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package."""
