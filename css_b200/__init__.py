"""css_b200 -- B200-native (sm_100a) representation-space hot path of CSS ("Space Engage", ICCV'23)."""
