"""oracle/ — TEST INFRASTRUCTURE ONLY: CPU restatement of the reference algorithms plus the build
recipe for the reference's own code (oracle/_ref/).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package."""
