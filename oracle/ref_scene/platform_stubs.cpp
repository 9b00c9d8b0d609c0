// TEST INFRASTRUCTURE. Minimal definitions of the platform/logging symbols the reference's
// frontend TUs link against (declared in Core/blitMemory.h:17-21, Core/DbLog/blitLogger.h:16-20,
// Core/Events/blitTimeManager.h:29-31, Core/DbLog/blitAssert.h:15, Core/blitzenEngine.h:145),
// so that the reference's camera / mesh / render-object code can run headless.
#include "Core/blitMemory.h"
#include "Core/Events/blitTimeManager.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace BlitzenPlatform
{
    void* PlatformMalloc(size_t size, uint8_t) { return malloc(size); }
    void PlatformFree(void* p, uint8_t) { free(p); }
    void* PlatformMemZero(void* p, size_t n) { return memset(p, 0, n); }
    void* PlatformMemCopy(void* d, void* s, size_t n) { return memcpy(d, s, n); }
    void* PlatformMemSet(void* d, int32_t v, size_t n) { return memset(d, v, n); }
    void PlatformConsoleWrite(const char* m, uint8_t) { if (getenv("REFSCENE_VERBOSE")) fputs(m, stderr); }
    void PlatformConsoleError(const char* m, uint8_t) { fputs(m, stderr); }
    void PlatformLoggerFileWrite(const char*, uint8_t) {}
    void PlatformLoggerFileError(const char*, uint8_t) {}
    void PlatfrormSetupClock(BlitzenCore::WorldTimerManager*) {}
    double PlatformGetAbsoluteTime(double) { return 0.0; }
}
namespace BlitzenCore
{
    void ReportAssertionFailure(const char* e, const char* m, const char* f, int32_t l)
    {
        fprintf(stderr, "ASSERT %s (%s) %s:%d\n", e, m, f, l);
        abort();
    }
    void ShutdownLogging(size_t, size_t*) {}
    WorldTimerManager::WorldTimerManager() : m_startTime(0), m_elapsedTime(0), m_previousTime(0), m_deltaTime(0), m_clockFrequency(1) {}
}
