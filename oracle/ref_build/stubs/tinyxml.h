// Include-path stub for <tinyxml.h> (oracle/_ref build only): forward declarations.
#ifndef GCM_ORACLE_STUB_TINYXML_H
#define GCM_ORACLE_STUB_TINYXML_H
class TiXmlElement;
class TiXmlDocument;
class TiXmlNode;
#endif
