"""Re-export of the seeded input generators (inputs only — not a checker)."""
from workoutdetector_b200.utils.synth import synth_clips_u8, synth_frames_u8, synth_video_u8  # noqa: F401
