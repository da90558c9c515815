"""harness -- bench and test support around the product package (trgt_b200/): the synthetic HiFi workload
generator (csrc/synth.cpp), the stand-in for the reference's host genotyper between the GPU phases
(csrc/hostglue.cpp), the workload containers (workload.py) and the phase pipeline bench.py times
(pipeline.py).  None of it is part of the drop-in: libtrgt_b200.so and trgt_b200/ never import it, and the
reference arm of bench.py loads only this package's libtrgt_harness.so and the oracle."""
