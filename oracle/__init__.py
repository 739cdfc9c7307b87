"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference algorithms (the parity oracle).

Nothing under selavi_b200/ may import this package.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or the timed CPU arm.
"""
