"""Stub of the un-vendored third-party package `allennlp==0.8.4` (reference requirements.txt:1).

TEST INFRASTRUCTURE ONLY. Just enough surface for the *unmodified* reference modules under
/root/reference to import in this container so that golden vectors can be generated
(see oracle/gen_golden.py). Never imported by the product path.
"""
