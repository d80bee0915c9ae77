// bcd::ParametersIO (reference: /root/reference/src/io/ParametersIO.cpp:36-162): the same keys, the same relative-path
// convention, the same "missing key keeps the caller's value" rule. The reference's write() falls off the end of a
// bool function (:162); here it returns whether the file was written.
#include "ParametersIO.h"
#include "Utils.h"

#include <cctype>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

namespace bcd
{

	namespace
	{
		// One level of JSON: an object whose values are strings, numbers, booleans or null; nested containers are
		// accepted and skipped.
		struct JsonValue
		{
			enum Kind { STRING, NUMBER, BOOL, OTHER } kind;
			std::string s;
			double d;
			bool b;
		};

		struct JsonParser
		{
			const std::string& t;
			size_t i;
			explicit JsonParser(const std::string& text) : t(text), i(0) {}
			void ws() { while(i < t.size() && isspace((unsigned char)t[i])) ++i; }
			bool lit(const char* w)
			{
				const size_t n = strlen(w);
				if(t.compare(i, n, w) != 0) return false;
				i += n;
				return true;
			}
			bool string(std::string& out)
			{
				if(i >= t.size() || t[i] != '"') return false;
				++i;
				out.clear();
				while(i < t.size() && t[i] != '"')
				{
					char c = t[i++];
					if(c == '\\')
					{
						if(i >= t.size()) return false;
						const char e = t[i++];
						switch(e)
						{
							case 'n': c = '\n'; break;
							case 't': c = '\t'; break;
							case 'r': c = '\r'; break;
							case 'b': c = '\b'; break;
							case 'f': c = '\f'; break;
							case 'u':
							{
								if(i + 4 > t.size()) return false;
								const unsigned cp = (unsigned)strtoul(t.substr(i, 4).c_str(), nullptr, 16);
								i += 4;
								if(cp < 0x80) out.push_back((char)cp);
								else if(cp < 0x800) { out.push_back((char)(0xc0 | cp >> 6)); out.push_back((char)(0x80 | (cp & 0x3f))); }
								else { out.push_back((char)(0xe0 | cp >> 12)); out.push_back((char)(0x80 | ((cp >> 6) & 0x3f))); out.push_back((char)(0x80 | (cp & 0x3f))); }
								continue;
							}
							default: c = e; break; // \" \\ \/
						}
					}
					out.push_back(c);
				}
				if(i >= t.size()) return false;
				++i;
				return true;
			}
			bool skipContainer(char open, char close)
			{
				int depth = 0;
				std::string dummy;
				while(i < t.size())
				{
					if(t[i] == '"') { if(!string(dummy)) return false; continue; }
					if(t[i] == open) ++depth;
					else if(t[i] == close && --depth == 0) { ++i; return true; }
					++i;
				}
				return false;
			}
			bool value(JsonValue& v)
			{
				ws();
				if(i >= t.size()) return false;
				const char c = t[i];
				if(c == '"') { v.kind = JsonValue::STRING; return string(v.s); }
				if(c == '{') { v.kind = JsonValue::OTHER; return skipContainer('{', '}'); }
				if(c == '[') { v.kind = JsonValue::OTHER; return skipContainer('[', ']'); }
				if(lit("true")) { v.kind = JsonValue::BOOL; v.b = true; return true; }
				if(lit("false")) { v.kind = JsonValue::BOOL; v.b = false; return true; }
				if(lit("null")) { v.kind = JsonValue::OTHER; return true; }
				char* end = nullptr;
				v.d = strtod(t.c_str() + i, &end);
				if(end == t.c_str() + i) return false;
				i = (size_t)(end - t.c_str());
				v.kind = JsonValue::NUMBER;
				return true;
			}
			bool object(std::map<std::string, JsonValue>& out)
			{
				ws();
				if(i >= t.size() || t[i] != '{') return false;
				++i;
				ws();
				if(i < t.size() && t[i] == '}') { ++i; return true; }
				while(true)
				{
					ws();
					std::string key;
					if(!string(key)) return false;
					ws();
					if(i >= t.size() || t[i] != ':') return false;
					++i;
					JsonValue v;
					if(!value(v)) return false;
					out[key] = v;
					ws();
					if(i < t.size() && t[i] == ',') { ++i; continue; }
					if(i < t.size() && t[i] == '}') { ++i; return true; }
					return false;
				}
			}
		};

		typedef std::map<std::string, JsonValue> JsonObject;

		void getString(const JsonObject& o, const char* key, const std::string& folder, std::string& dst)
		{
			JsonObject::const_iterator it = o.find(key);
			if(it != o.end() && it->second.kind == JsonValue::STRING)
				dst = folder + it->second.s;
		}
		template <typename T> void getNumber(const JsonObject& o, const char* key, T& dst)
		{
			JsonObject::const_iterator it = o.find(key);
			if(it == o.end()) return;
			if(it->second.kind == JsonValue::NUMBER) dst = static_cast<T>(it->second.d);
			else if(it->second.kind == JsonValue::BOOL) dst = static_cast<T>(it->second.b ? 1 : 0);
		}
		void getBool(const JsonObject& o, const char* key, bool& dst)
		{
			JsonObject::const_iterator it = o.find(key);
			if(it == o.end()) return;
			if(it->second.kind == JsonValue::BOOL) dst = it->second.b;
			else if(it->second.kind == JsonValue::NUMBER) dst = it->second.d != 0.0;
		}

		std::string quoted(const std::string& s)
		{
			std::string q = "\"";
			for(size_t k = 0; k < s.size(); ++k)
			{
				const unsigned char c = (unsigned char)s[k];
				if(c == '"' || c == '\\') { q.push_back('\\'); q.push_back((char)c); }
				else if(c == '\n') q += "\\n";
				else if(c == '\t') q += "\\t";
				else if(c == '\r') q += "\\r";
				else if(c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); q += b; }
				else q.push_back((char)c);
			}
			return q + "\"";
		}
		std::string number(int v)
		{
			char b[32];
			snprintf(b, sizeof b, "%d", v);
			return b;
		}

		std::string number(float v)
		{
			// shortest text that reads back as the same float; a float field always reads as a JSON real ("1.0", like
			// nlohmann::json writes it), so that a strict reader sees the same types the reference's files have
			char b[64];
			for(int prec = 1; prec <= 17; ++prec)
			{
				snprintf(b, sizeof b, "%.*g", prec, (double)v);
				if((float)strtod(b, nullptr) == v) break;
			}
			std::string s = b;
			if(s.find_first_of(".eEn") == std::string::npos) s += ".0";
			return s;
		}
	}

	bool ParametersIO::load(PipelineParameters& o_rParams, const std::string& i_rFilePath, PipelineParametersSelector i_selector)
	{
		if(i_rFilePath.empty())
		{
			std::cerr << "Couldn't load parameters: empty file name" << std::endl;
			return false;
		}
		std::ifstream file(i_rFilePath.c_str());
		if(!file)
		{
			std::cerr << "Error: couldn't open file '" << i_rFilePath << "'" << std::endl;
			return false;
		}
		std::stringstream ss;
		ss << file.rdbuf();
		const std::string text = ss.str();
		JsonObject o;
		JsonParser parser(text);
		if(!parser.object(o))
		{
			std::cerr << "Error: '" << i_rFilePath << "' is not a JSON object (near character " << parser.i << ")" << std::endl;
			return false;
		}
		const std::string folder = Utils::extractFolderPath(i_rFilePath);
		if(i_selector.m_inputFileNames)
		{
			getString(o, "inputColorFile", folder, o_rParams.m_inputFileNames.m_colors);
			getString(o, "inputHistoFile", folder, o_rParams.m_inputFileNames.m_histograms);
			getString(o, "inputCovarFile", folder, o_rParams.m_inputFileNames.m_covariances);
		}
		if(i_selector.m_prefilteringParameters)
		{
			getBool(o, "performSpikeRemovalPrefiltering", o_rParams.m_prefilteringParameters.m_performSpikeRemoval);
			getNumber(o, "spikeRemovalThresholdStDevFactor", o_rParams.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor);
		}
		if(i_selector.m_denoiserParameters)
		{
			getNumber(o, "nbOfScales", o_rParams.m_denoiserParameters.m_nbOfScales);
			DenoiserParameters& p = o_rParams.m_denoiserParameters.m_monoscaleParameters;
			getNumber(o, "histoDistanceThreshold", p.m_histogramDistanceThreshold);
			getBool(o, "useCuda", p.m_useCuda);
			getNumber(o, "nbOfCores", p.m_nbOfCores);
			getNumber(o, "patchRadius", p.m_patchRadius);
			getNumber(o, "searchWindowRadius", p.m_searchWindowRadius);
			getBool(o, "randomPixelOrder", p.m_useRandomPixelOrder);
			getNumber(o, "markedPixelsSkippingProbability", p.m_markedPixelsSkippingProbability);
			getNumber(o, "minEigenValue", p.m_minEigenValue);
		}
		return true;
	}

	bool ParametersIO::write(const PipelineParameters& i_rParams, const std::string& i_rFilePath, PipelineParametersSelector i_selector)
	{
		if(i_rFilePath.empty())
		{
			std::cerr << "Couldn't save parameters: empty file name" << std::endl;
			return false;
		}
		std::ofstream file(i_rFilePath.c_str());
		if(!file)
		{
			std::cerr << "Error: couldn't write file '" << i_rFilePath << "'" << std::endl;
			return false;
		}
		const std::string folder = Utils::extractFolderPath(i_rFilePath);
		std::map<std::string, std::string> kv; // sorted keys, 4-space indentation: the layout nlohmann's dump(4) produces
		if(i_selector.m_inputFileNames)
		{
			const InputFileNames& n = i_rParams.m_inputFileNames;
			kv["inputColorFile"] = quoted(Utils::getRelativePathFromFolder(n.m_colors, folder));
			kv["inputHistoFile"] = quoted(Utils::getRelativePathFromFolder(n.m_histograms, folder));
			kv["inputCovarFile"] = quoted(Utils::getRelativePathFromFolder(n.m_covariances, folder));
		}
		if(i_selector.m_prefilteringParameters)
		{
			kv["performSpikeRemovalPrefiltering"] = i_rParams.m_prefilteringParameters.m_performSpikeRemoval ? "true" : "false";
			kv["spikeRemovalThresholdStDevFactor"] = number(i_rParams.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor);
		}
		if(i_selector.m_denoiserParameters)
		{
			const DenoiserParameters& p = i_rParams.m_denoiserParameters.m_monoscaleParameters;
			kv["nbOfScales"] = number(i_rParams.m_denoiserParameters.m_nbOfScales);
			kv["histoDistanceThreshold"] = number(p.m_histogramDistanceThreshold);
			kv["useCuda"] = p.m_useCuda ? "true" : "false";
			kv["nbOfCores"] = number(p.m_nbOfCores);
			kv["patchRadius"] = number(p.m_patchRadius);
			kv["searchWindowRadius"] = number(p.m_searchWindowRadius);
			kv["randomPixelOrder"] = p.m_useRandomPixelOrder ? "true" : "false";
			kv["markedPixelsSkippingProbability"] = number(p.m_markedPixelsSkippingProbability);
			kv["minEigenValue"] = number(p.m_minEigenValue);
		}
		file << "{";
		bool firstKey = true;
		for(std::map<std::string, std::string>::const_iterator it = kv.begin(); it != kv.end(); ++it, firstKey = false)
			file << (firstKey ? "\n" : ",\n") << "    " << quoted(it->first) << ": " << it->second;
		file << (kv.empty() ? "}" : "\n}");
		file.flush();
		return static_cast<bool>(file);
	}

} // namespace bcd
