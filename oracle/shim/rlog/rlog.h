// oracle shim (test infrastructure only): rlog logging (F/RLogInterface.h:26-42) -> no-op.
#pragma once
namespace rlog { class RLogChannel; class StdioNode; }
#define LOGID 0
#define _rMessage(...) ((void)0)
#define RLOG_CHANNEL(x) ((rlog::RLogChannel*)0)
#define rDebug(...) ((void)0)
#define rInfo(...) ((void)0)
#define rWarning(...) ((void)0)
#define rError(...) ((void)0)
