"""oracle/ — TEST INFRASTRUCTURE.  CPU restatements of the reference's algorithms.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under ape_b200/ does; the product path has no CPU fallback."""
