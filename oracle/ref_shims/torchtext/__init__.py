"""Stub of `torchtext` (imported at reference updown_captioner.py:9). With an empty `stoi`
every vocabulary word takes the reference's own OOV initialisation 2*randn-1 (:197,:209,:215)."""
