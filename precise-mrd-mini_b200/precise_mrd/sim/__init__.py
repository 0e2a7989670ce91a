from .contamination import ContaminationSimulator, run_contamination_stress_test  # noqa: F401
