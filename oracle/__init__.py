"""TEST INFRASTRUCTURE ONLY — the CPU oracle. Import from tests/, __graft_entry__.smoke() and bench.py's CPU legs, nowhere else."""
