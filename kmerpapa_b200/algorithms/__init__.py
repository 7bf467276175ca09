"""Drop-in replacements of the two reference estimators the command line calls on the hot path."""
