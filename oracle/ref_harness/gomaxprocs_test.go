package framework

import goruntime "runtime"

// kept apart so that the harness file does not collide with the import alias the reference's own test file uses
func goruntimeGOMAXPROCS() int { return goruntime.GOMAXPROCS(0) }
