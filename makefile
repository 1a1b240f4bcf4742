#####################################################
# The reference's make targets (lingxuez/URSM makefile:15-41) on the sm_100a path.
#
#   make demo      build, simulate the demo inputs on the device, fit, check the recovery
#   make run       the reference's `make run` command line, unchanged
#   make check     numeric equivalent of demo_plots.py (correlation of est. vs true A, W)
#   make simulate  re-draw demo/demo_data/*.csv (demo_simulate_data.py's distributions)
#####################################################
.PHONY: build run check plot demo simulate clean test

PYTHON ?= python

demo: build simulate run check clean

build:
	$(PYTHON) -m ursm_b200.build

simulate:
	$(PYTHON) demo/demo_simulate_data.py

run:
	@mkdir -p demo/demo_out
	$(PYTHON) scUnif.py \
	-K 3 \
	-sc demo/demo_data/demo_single_cell_rnaseq_counts.csv \
	-ctype demo/demo_data/demo_single_cell_types.csv \
	-bk demo/demo_data/demo_bulk_rnaseq_counts.csv \
	-outdir demo/demo_out \
	-log demo/demo_out/demo_logging.log \
	-burnin 10 -sample 20 \
	-EM_maxiter 10

check:
	$(PYTHON) demo/demo_check.py

plot: check

test:
	$(PYTHON) -m pytest tests -q -m "not gpu"

clean:
	@rm -f *.pyc
	@rm -rf __pycache__ */__pycache__
