"""CPU oracle of the rl-rep update step.  TEST INFRASTRUCTURE ONLY.

Nothing under rlrep_b200/ may import this package; only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py do, and only as the checker or the timed baseline.
"""
