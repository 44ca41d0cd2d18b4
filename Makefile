# Convenience targets; the real build logic lives in native/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
PY ?= python

.PHONY: native test test-gpu bench sass clean
native:
	$(PY) native/build.py
test: native
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu: native
	$(PY) -m pytest tests -x -q -m gpu
bench: native
	$(PY) bench.py
sass: native
	./profiles/collect_sass.sh
clean:
	rm -rf build batch_shipyard_b200/_native
