from .monitor import Monitor, ResultsWriter, get_monitor_files, load_results  # noqa: F401
