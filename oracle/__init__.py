"""CPU checkers for the tracking hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; nothing under maveric-slam_b200/ does.
"""
