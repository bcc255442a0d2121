"""CPU checkers for the CSGN hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; nothing under csgn_b200/ does.
"""
