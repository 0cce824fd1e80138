"""oracle/ -- CPU restatement of the reference's hot-path algorithms.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under prism_b200/ does (tests/test_boundary.py checks that).
"""
