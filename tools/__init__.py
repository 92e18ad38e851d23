"""Fixture tooling shared by tests, bench.py and __graft_entry__.smoke(): decoder hyper-parameter
dataclasses and the seeded synthetic-checkpoint generator.  Neither product code (the library never
needs it) nor oracle code (it computes nothing the decoder computes)."""
