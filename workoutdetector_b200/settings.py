"""Path constants of the reference (workoutdetector/settings/global_settings.py:3-10), env-overridable."""
import os

PROJ_ROOT = os.environ.get('PROJ_ROOT', '/work')
DATA_ROOT = os.path.expanduser('~/data')
REPCOUNT_ANNO_PATH = os.path.join(PROJ_ROOT, 'datasets/RepCount/annotation.csv')
