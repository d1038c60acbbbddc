"""CPU oracles (test infrastructure). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this."""
