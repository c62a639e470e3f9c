/* oracle/ref_shim_inc/portaudio.h — TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in so that the reference's server headers (src/server/client.h -> audio/audio.h) parse when
 * src/server/stream.c is compiled for the oracle.  No audio code is compiled or called. */
#pragma once
typedef void PaStream;
typedef int PaError;
typedef int PaDeviceIndex;
typedef int PaHostApiIndex;
typedef double PaTime;
typedef unsigned long PaStreamCallbackFlags;
typedef unsigned long PaSampleFormat;
typedef unsigned long PaStreamFlags;
typedef struct PaStreamCallbackTimeInfo {
  PaTime inputBufferAdcTime, currentTime, outputBufferDacTime;
} PaStreamCallbackTimeInfo;
typedef struct PaStreamParameters {
  PaDeviceIndex device;
  int channelCount;
  PaSampleFormat sampleFormat;
  PaTime suggestedLatency;
  void *hostApiSpecificStreamInfo;
} PaStreamParameters;
typedef struct PaDeviceInfo {
  int structVersion;
  const char *name;
  PaHostApiIndex hostApi;
  int maxInputChannels, maxOutputChannels;
  PaTime defaultLowInputLatency, defaultLowOutputLatency, defaultHighInputLatency, defaultHighOutputLatency;
  double defaultSampleRate;
} PaDeviceInfo;
typedef int PaStreamCallback(const void *, void *, unsigned long, const PaStreamCallbackTimeInfo *,
                             PaStreamCallbackFlags, void *);
#define paNoError 0
#define paFloat32 1
#define paContinue 0
