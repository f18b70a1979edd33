"""TEST INFRASTRUCTURE ONLY: the CPU oracle for the MapRead hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (lra_b200/) never does.
"""
