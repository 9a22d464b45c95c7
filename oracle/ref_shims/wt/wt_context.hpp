// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the application context (include/wt/wt_context.hpp: paths, thread pool, flags)
// is only named, never used, by the headers compiled there.
#pragma once
namespace wt { struct wt_context_t {}; }
